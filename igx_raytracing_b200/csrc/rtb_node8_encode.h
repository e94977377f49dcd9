// rtb_node8_encode.h — the Node8 box encoding (rtb_types.h), shared by the host builder (rtb_bvh.cpp) and the device
// refit (rtb_refit.cu) so that both round the same way: every stored child box CONTAINS the box it was given.
//
//   node8Grid   origin = lo corner of the node box; per-axis power-of-two step with extent / step <= 254
//   node8Child  child box -> bf16 grid coordinates, lo rounded towards -inf and hi towards +inf after a 0.02-step margin;
//               slots in the upper half of a word are read by the traversal together with the 16 bits below them, so an
//               upper-half lo plane is stored one extra step lower (rtb_trace8.cuh, testPair)
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "rtb_types.h"

#ifdef __CUDACC__
#define RTB_ENC_HD __host__ __device__ inline
#else
#define RTB_ENC_HD inline
#endif

namespace rtb {

struct Box6 { float lo[3], hi[3]; };

// smallest e with ext / 2^e <= 254 (integer arithmetic on the exponent: identical on host and device), clamped to +-100
RTB_ENC_HD int node8Exponent(double ext) {
    if (!(ext > 0.0)) return -100;
    int x;
    const double m = frexp(ext / 254.0, &x);   // ext / 254 = m * 2^x, m in [0.5, 1)
    int e = m == 0.5 ? x - 1 : x;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    while (ext / ldexp(1.0, e) > 254.0 && e < 100) ++e;
    return e;
}

RTB_ENC_HD void node8Grid(const Box6& nb, Node8& out, double step[3]) {
    for (int a = 0; a < 3; ++a) {
        out.p[a] = nb.lo[a];
        const int e = node8Exponent((double)nb.hi[a] - (double)nb.lo[a]);
        step[a] = ldexp(1.0, e);
        out.e[a] = (uint8_t)(e + 127);
    }
}

RTB_ENC_HD double node8Bf16Value(uint32_t topBits) { float f; memcpy(&f, &topBits, 4); return (double)f; }

// bf16 bits of a non-negative grid coordinate, rounded towards -inf / +inf
RTB_ENC_HD uint32_t node8Bf16Down(double g) {
    g = g > 0.0 ? g : 0.0;
    const float f = (float)g; uint32_t b; memcpy(&b, &f, 4);
    uint32_t t = b & 0xFFFF0000u;                                 // truncation = towards zero
    if (node8Bf16Value(t) > g && t >= 0x10000u) t -= 0x10000u;    // (float)g rounded up onto a bf16 value
    return t >> 16;
}
RTB_ENC_HD uint32_t node8Bf16Up(double g) {
    g = g > 0.0 ? g : 0.0;
    const float f = (float)g; uint32_t b; memcpy(&b, &f, 4);
    uint32_t t = b & 0xFFFF0000u;
    if (node8Bf16Value(t) < g) t += 0x10000u;
    return t >> 16;
}

// box == nullptr: an empty slot (an inverted box; the valid mask removes its bits anyway).  `out.planes` must start zeroed.
RTB_ENC_HD void node8Child(Node8& out, int slot, const Box6* box, const double step[3]) {
    const bool upper = (slot & 1) == 0;   // even slots live in the upper half of their word and are read without decoding
    const int word = slot >> 1, shift = upper ? 16 : 0;
    if (!box) {
        for (int a = 0; a < 3; ++a) out.planes[a][0][word] |= 0x4380u << shift;   // lo = 256.0
        return;
    }
    for (int a = 0; a < 3; ++a) {
        const double lo = ((double)box->lo[a] - (double)out.p[a]) / step[a] - 0.02, hi = ((double)box->hi[a] - (double)out.p[a]) / step[a] + 0.02;
        uint32_t ql = node8Bf16Down(lo);
        const uint32_t qh = node8Bf16Up(hi);
        if (upper && ql > 0) ql -= 1;
        out.planes[a][0][word] |= ql << shift;
        out.planes[a][1][word] |= qh << shift;
    }
}

// padded box of one reference-layout triangle, as the builder boxes it (rtb_bvh.cpp, buildBinary)
RTB_ENC_HD Box6 node8TriangleBox(const TriangleRec& t, float maxAbs, float pad) {
    Box6 b;   // fminf / fmaxf skip NaN operands, as std::min(acc, x) does in the builder
    for (int a = 0; a < 3; ++a) {
        b.lo[a] = fminf(fminf(t.p0[a], t.p1[a]), t.p2[a]) - pad;
        b.hi[a] = fmaxf(fmaxf(t.p0[a], t.p1[a]), t.p2[a]) + pad;
        if (!(b.lo[a] <= b.hi[a])) { b.lo[a] = -maxAbs - pad; b.hi[a] = maxAbs + pad; }   // NaN vertex: keep it reachable
    }
    return b;
}

RTB_ENC_HD float node8Pad(float maxAbs) { const float p = maxAbs * 3.814697265625e-6f; /* 2^-18 */ return p > 1e-30f ? p : 1e-30f; }

}  // namespace rtb
