// rtb_crmath.h — sin / cos of a binary32 argument evaluated in binary64 and rounded once (decree D4 of the numerics
// contract), without the slow path of a general binary64 sin.
//
// The shader RNG (SH/rand_util.glsl:115-129) takes sin of arguments up to ~3e14 (`p * 1103515245 + 12345` dotted with
// (12.9898, 78.233)); every call lands in the Payne–Hanek path of CUDA's sin(double) (|x| > 105615), which dominates
// k_raygen / k_shadowgen / k_shade.  A binary32 argument below 2^50 does not need it: with k = rint(x * 2/pi) < 2^50,
//     r = fma(-k, PIO2_HI, x);  r = fma(-k, PIO2_LO, r)
// leaves |error| < 2^-52 (the products are exact inside the fma, PIO2_HI + PIO2_LO carries 106 bits of pi/2, and the
// tail k * 2^-107 stays below 2^-57), after which the usual degree-13 / degree-14 minimax kernels on [-pi/4, pi/4]
// (coefficients as published in FreeBSD msun k_sin.c / k_cos.c) give sin and cos to under one binary64 ulp.  The
// binary32 rounding of that equals the rounding of the exact value except within ~2^-28 of a rounding boundary —
// the same budget the parity tests already carry for CUDA-vs-glibc binary64 differences (tests/test_gpu_parity.py).
// Arguments of 2^50 and above, infinities and NaN go to the general routine.
//
// Host-compilable (tests/cpp/crmath_check.cpp compares against glibc on 2^28 arguments).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define RTB_HD __host__ __device__ __forceinline__
#else
#define RTB_HD inline
#endif

namespace rtb {

RTB_HD double crKernelSin(double r) {
    const double z = r * r;
    double p = 1.58969099521155010221e-10;
    p = fma(p, z, -2.50507602534068634195e-08);
    p = fma(p, z, 2.75573137070700676789e-06);
    p = fma(p, z, -1.98412698298579493134e-04);
    p = fma(p, z, 8.33333333332248946124e-03);
    p = fma(p, z, -1.66666666666666324348e-01);
    return fma(z * r, p, r);
}

RTB_HD double crKernelCos(double r) {
    const double z = r * r;
    double p = -1.13596475577881948265e-11;
    p = fma(p, z, 2.08757232129817482790e-09);
    p = fma(p, z, -2.75573143513906633035e-07);
    p = fma(p, z, 2.48015872894767294178e-05);
    p = fma(p, z, -1.38888888888741095749e-03);
    p = fma(p, z, 4.16666666666666019037e-02);
    // 1 - z/2 + z^2 p, with the rounding error of (1 - z/2) carried along
    const double hz = 0.5 * z, w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + z * z * p);
}

// quadrant q (mod 4) and remainder r in about [-pi/4, pi/4] of a finite |x| < 2^50
RTB_HD double crReduce(double x, int& q) {
    const double k = rint(x * 6.36619772367581382433e-01);
    q = (int)((long long)k & 3);
    if (k == 0.0) return x;
    double r = fma(-k, 1.57079632679489655800e+00, x);
    r = fma(-k, 6.12323399573676603587e-17, r);
    return r;
}

RTB_HD float cr_sin_f(float xf) {
    const double x = (double)xf;
    if (!(fabs(x) < 1125899906842624.0)) return (float)sin(x);
    if (x == 0.0) return xf;   // sin(-0) = -0
    int q;
    const double r = crReduce(x, q);
    const double v = (q & 1) ? crKernelCos(r) : crKernelSin(r);
    return (float)((q & 2) ? -v : v);
}

RTB_HD float cr_cos_f(float xf) {
    const double x = (double)xf;
    if (!(fabs(x) < 1125899906842624.0)) return (float)cos(x);
    int q;
    const double r = crReduce(x, q);
    const double v = (q & 1) ? crKernelSin(r) : crKernelCos(r);
    return (float)(((q + 1) & 2) ? -v : v);
}

// pow(x, 5) as the shaders use it (SH/light.glsl:58-60: pow(1 - NdotV, 5.0)): three binary64 products, relative error
// < 2^-51, one rounding to binary32.  Matches pow() for every sign, zero, infinity and NaN (odd integer exponent).
RTB_HD float cr_pow5_f(float xf) {
    const double x = (double)xf, x2 = x * x;
    return (float)(x2 * x2 * x);
}

}  // namespace rtb
