// crmath_check.cpp — host check of igx_raytracing_b200/csrc/rtb_crmath.h against glibc's binary64 functions rounded once
// (what the oracle computes, oracle/oracle.cpp D4).  Usage: crmath_check [stride]   (stride over binary32 bit patterns)
// Prints the number of arguments whose binary32 result differs; exit status 1 when the mismatch rate exceeds 1e-7.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "rtb_crmath.h"

static float fromBits(uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; }
static uint32_t bitsOf(float f) { uint32_t b; std::memcpy(&b, &f, 4); return b; }

int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 37u;
    uint64_t n = 0, badSin = 0, badCos = 0, badPow = 0;
    // every stride-th non-negative binary32 up to and beyond the 2^50 switch-over (0x58800000), both signs
    #pragma omp parallel for reduction(+ : n, badSin, badCos, badPow) schedule(static)
    for (int64_t i = 0; i < (int64_t)(0x5A000000u / stride); ++i) {
        const uint32_t b = (uint32_t)i * stride;
        for (int s = 0; s < 2; ++s) {
            const float x = fromBits(b | (s ? 0x80000000u : 0u));
            ++n;
            if (bitsOf(rtb::cr_sin_f(x)) != bitsOf((float)std::sin((double)x))) ++badSin;
            if (bitsOf(rtb::cr_cos_f(x)) != bitsOf((float)std::cos((double)x))) ++badCos;
            if (bitsOf(rtb::cr_pow5_f(x)) != bitsOf((float)std::pow((double)x, 5.0))) ++badPow;
        }
    }
    // the RNG's own arguments: dot(p, k) and dot(p * 1103515245 + 12345, k) for pixel-like p (SH/rand_util.glsl:115-129)
    uint64_t m = 0, badRng = 0;
    #pragma omp parallel for reduction(+ : m, badRng) schedule(static)
    for (int y = 0; y < 2160; y += 3) {
        for (int x = 0; x < 3840; x += 5) {
            const float px = (float)x + 0.37f, py = (float)y + 0.81f;
            const float a0 = px * 12.9898f + py * 78.233f;
            const float qx = px * 1103515245.0f + 12345.0f, qy = py * 1103515245.0f + 12345.0f;
            const float a1 = qx * 12.9898f + qy * 78.233f;
            m += 2;
            if (bitsOf(rtb::cr_sin_f(a0)) != bitsOf((float)std::sin((double)a0))) ++badRng;
            if (bitsOf(rtb::cr_sin_f(a1)) != bitsOf((float)std::sin((double)a1))) ++badRng;
        }
    }
    // specials
    const float sp[] = {0.0f, -0.0f, INFINITY, -INFINITY, NAN, 1e-45f, 3.4028235e38f, 1125899906842624.0f, 1125899839733760.0f};
    int badSpecial = 0;
    for (float x : sp) {
        const float a = rtb::cr_sin_f(x), b = (float)std::sin((double)x), c = rtb::cr_cos_f(x), d = (float)std::cos((double)x);
        const float e = rtb::cr_pow5_f(x), f = (float)std::pow((double)x, 5.0);
        if (!((a != a && b != b) || bitsOf(a) == bitsOf(b))) ++badSpecial;
        if (!((c != c && d != d) || bitsOf(c) == bitsOf(d))) ++badSpecial;
        if (!((e != e && f != f) || bitsOf(e) == bitsOf(f))) ++badSpecial;
    }
    std::printf("args %llu  sin %llu  cos %llu  pow5 %llu  |  rng args %llu  sin %llu  |  specials bad %d\n", (unsigned long long)n,
                (unsigned long long)badSin, (unsigned long long)badCos, (unsigned long long)badPow, (unsigned long long)m,
                (unsigned long long)badRng, badSpecial);
    const double rate = (double)(badSin + badCos + badPow) / (3.0 * (double)n);
    return (rate > 1e-7 || badRng > 2 || badSpecial) ? 1 : 0;
}
