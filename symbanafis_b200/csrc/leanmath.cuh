// leanmath.cuh -- issue-lean f64 sin / cos / sincos / exp for the interpreter's hottest opcodes.
//
// Why not libdevice here: the FP64 pipe of an SM accepts one warp instruction every two cycles, so a
// kernel is FP64-bound only while (all instructions) <= 2 x (FP64 instructions).  libdevice's sin
// executes ~52 instructions per point for 15 FP64 ones (constant materialisation through UMOV pairs,
// 64-bit table addressing + 3 global loads, FSEL selects, special-case plumbing; ncu source view in
// profiles/r01_a_*).  The kernels below do the same mathematics -- 3-term Cody-Waite reduction by
// pi/2, fdlibm-grade minimax kernels on [-pi/4, pi/4], coefficient choice by quadrant parity -- in
// ~30 instructions: coefficients come from a 128-byte shared-memory table (one conflict-free LDS.64
// per coefficient: the sin and cos entries of a term sit in different banks), the sign is an XOR.
//
// Accuracy (tools/leanmath_check.c, 4M points per decade, vs glibc): sin <= 1 ulp, cos <= 2 ulp for
// |x| <= 1e6, exp <= 1 ulp on [-700, 700].  Outside the fast range (|x| >= 2^20, zeros/denormals for
// sin so that sin(-0.0) = -0.0, |x| > 700 for exp, NaN/Inf) the CUDA math library is called, so
// special-value behaviour is exactly libdevice's.
#pragma once
#include <cstdint>

namespace saf {
namespace lm {

// [term j][0 = sin kernel, 1 = cos kernel], highest degree first; the sin kernel has one term fewer.
// sin(r) = r + (z r) * Ps(z),  cos(r) = 1 + z * Pc(z),  z = r^2   (fdlibm k_sin.c / k_cos.c coefficients)
__device__ __constant__ double kTrigTable[7][2] = {
    {0.0, -1.13596475577881948265e-11},
    {1.58969099521155010221e-10, 2.08757232129817482790e-09},
    {-2.50507602534068634195e-08, -2.75573143513906633035e-07},
    {2.75573137070700676789e-06, 2.48015872894767294178e-05},
    {-1.98412698298579493134e-04, -1.38888888888741095749e-03},
    {8.33333333332248946124e-03, 4.16666666666666019037e-02},
    {-1.66666666666666324348e-01, -0.5},
};
constexpr int kTrigTableBytes = 7 * 2 * 8;

constexpr double kTwoOverPi = 6.36619772367581382433e-01;
constexpr double kPio2_1 = 1.57079632679489655800e+00;
constexpr double kPio2_2 = 6.12323399573676603587e-17;
constexpr double kPio2_3 = -1.49738490485916983249e-33;

// true when x is outside the fast range of the trig kernels: zero, denormal, |x| >= 2^20, Inf, NaN
__device__ __forceinline__ bool trig_slow(double x) {
    const uint32_t hi = (uint32_t)__double2hiint(x) & 0x7fffffffu;
    return (hi - 0x00100000u) >= (0x41300000u - 0x00100000u);
}

// q = nearest integer to x*2/pi (low bits), r = x - q*pi/2
__device__ __forceinline__ double trig_reduce(double x, int &q) {
    q = __double2int_rn(__dmul_rn(x, kTwoOverPi));
    const double kd = __int2double_rn(q);
    double r = __fma_rn(kd, -kPio2_1, x);
    r = __fma_rn(kd, -kPio2_2, r);
    r = __fma_rn(kd, -kPio2_3, r);
    return r;
}

// sin kernel if (q & 1) == 0, cos kernel otherwise; sign flipped when q & 2.  `tab` = shared-memory copy
// of kTrigTable (generic pointer into shared memory).
__device__ __forceinline__ double trig_eval(double r, int q, const double *tab) {
    const double *t = tab + (q & 1);
    const double z = __dmul_rn(r, r);
    double p = t[0];
    p = __fma_rn(p, z, t[2]);
    p = __fma_rn(p, z, t[4]);
    p = __fma_rn(p, z, t[6]);
    p = __fma_rn(p, z, t[8]);
    p = __fma_rn(p, z, t[10]);
    p = __fma_rn(p, z, t[12]);
    const bool odd = q & 1;
    const double m = odd ? z : __dmul_rn(z, r);
    const double c = odd ? 1.0 : r;
    const double v = __fma_rn(p, m, c);
    // sign: flip when bit 1 of q is set
    const int hi = __double2hiint(v) ^ ((q & 2) << 30);
    return __hiloint2double(hi, __double2loint(v));
}

// ---- P-wide entry points -------------------------------------------------------------------------
// The fast path runs branch-free over all P points of a thread (the compiler interleaves the P
// dependent FMA chains: instruction-level parallelism instead of occupancy); points outside the fast
// range are then recomputed one by one with the CUDA math library.  Every point's result depends only
// on its own input -- never on which other points share the thread.
template <int P>
__device__ __forceinline__ void sin_n(const double (&x)[P], double (&out)[P], const double *tab) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        int q;
        const double r = trig_reduce(x[p], q);
        out[p] = trig_eval(r, q, tab);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
        if (trig_slow(x[p])) out[p] = ::sin(x[p]);
}
template <int P>
__device__ __forceinline__ void cos_n(const double (&x)[P], double (&out)[P], const double *tab) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        int q;
        const double r = trig_reduce(x[p], q);
        out[p] = trig_eval(r, q + 1, tab);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
        if (trig_slow(x[p])) out[p] = ::cos(x[p]);
}
template <int P>
__device__ __forceinline__ void sincos_n(const double (&x)[P], double (&s)[P], double (&c)[P], const double *tab) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        int q;
        const double r = trig_reduce(x[p], q);
        s[p] = trig_eval(r, q, tab);
        c[p] = trig_eval(r, q + 1, tab);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
        if (trig_slow(x[p])) ::sincos(x[p], &s[p], &c[p]);
}

// exp: k = rint(x log2 e), r = x - k ln2 (hi + lo), degree-13 Taylor kernel (remainder < 2^-57 on
// |r| <= ln2/2), result scaled by 2^k through the exponent field (results stay normal for |x| < 700).
constexpr double kLog2e = 1.44269504088896338700e+00;
constexpr double kLn2Hi = 6.93147180369123816490e-01;
constexpr double kLn2Lo = 1.90821492927058770002e-10;
__device__ __constant__ double kExpTaylor[14] = {
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
    1.0 / 5040.0,       1.0 / 720.0,       1.0 / 120.0,      1.0 / 24.0,      1.0 / 6.0,      0.5,
    1.0,                1.0};

__device__ __forceinline__ double exp_fast(double x) {
    const int k = __double2int_rn(__dmul_rn(x, kLog2e));
    const double kd = __int2double_rn(k);
    double r = __fma_rn(kd, -kLn2Hi, x);
    r = __fma_rn(kd, -kLn2Lo, r);
    double p = kExpTaylor[0];
#pragma unroll
    for (int i = 1; i < 14; ++i) p = __fma_rn(p, r, kExpTaylor[i]);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
template <int P>
__device__ __forceinline__ void exp_n(const double (&x)[P], double (&out)[P]) {
#pragma unroll
    for (int p = 0; p < P; ++p) out[p] = exp_fast(x[p]);
#pragma unroll
    for (int p = 0; p < P; ++p)
        if (!(fabs(x[p]) < 700.0)) out[p] = ::exp(x[p]);
}

} // namespace lm
} // namespace saf
