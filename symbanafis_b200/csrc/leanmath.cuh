// leanmath.cuh -- issue-lean f64 sin / cos / sincos / exp for the interpreter's hottest opcodes.
//
// Why not libdevice here: the FP64 pipe of an SM accepts one warp instruction every two cycles, so a
// kernel is FP64-bound only while (all instructions) <= 2 x (FP64 instructions), and the integer/select
// pipe is just as narrow.  libdevice's sin executes ~52 instructions per point for 15 FP64 ones
// (constant materialisation, table addressing + 3 loads, FSEL selects, special-case plumbing; ncu source
// view in profiles/r01_a_*); a quadrant-selecting kernel needs two selects per coefficient.  The kernels
// below contain NO loads, selects or conversions -- FP64 instructions with constant-bank coefficients plus
// three integer instructions per point:
//
//   sin(x) = (-1)^k sin(r),      k = rint(x / pi),          r = x - k pi           |r| <= pi/2
//   cos(x) = (-1)^(k+1) sin(r),  k = rint(x / pi - 1/2),    r = x - (k + 1/2) pi
//   rint by the 1.5 * 2^52 trick (k sits in the low word of the biased sum), three-term Cody-Waite
//   reduction with FMAs (pi split into three POSITIVE parts so that -0 survives: sin(-0) = -0),
//   sin(r) = r + r z P(z), z = r^2, P = degree-7 fit of (sin r - r) / r^3 on |r| <= 1.0002 pi/2
//   (Chebyshev nodes, 60-digit arithmetic: tools/fit_leanmath.py; |P error| z < 1e-18),
//   sign of the result = sign of the (parity-flipped) reduced argument, so zeros keep their sign.
//   exp(x) = 2^k (1 + r + r^2 Q(r)),  k = rint(x log2 e),  r = x - k ln2 (hi + lo), Q degree 10.
//
// Accuracy (tools/leanmath_check.c replays the same operation sequences on the CPU with C99 fma, 4M points
// per range, against 80-bit libm): sin <= 1.96 ulp, cos <= 2.04 ulp for |x| < 2^20, exp <= 0.86 ulp on
// [-700, 700]; against glibc's double functions (what the reference calls): <= 2 / 2 / 1 ulp.  Zero,
// denormal and -0 arguments go through the fast path with the exact libm result.  Outside the fast range
// (|x| >= 2^20, |x| >= 700 for exp, Inf, NaN) the CUDA math library is called, so special-value
// behaviour there is exactly libdevice's.
#pragma once
#include <cstdint>

namespace saf {
namespace lm {

constexpr double kMagic = 6755399441055744.0; // 1.5 * 2^52 (low word zero: an instruction immediate)

// Coefficients live in the constant bank (an FP64 instruction takes a c[bank][offset] operand for free),
// not in the instruction stream: a 64-bit immediate would cost two UMOV per use.
__device__ __constant__ double kTrig[12] = {
    3.18309886183790691e-01,  // [0] 1/pi
    3.14159265358979312e+00,  // [1] pi, part A
    1.22464679914735296e-16,  // [2] pi, part B -- rounded DOWN: every part of the split is positive
    2.16571334784382806e-32,  // [3] pi, part C
    -1.66666666666666657e-01, 8.33333333333331587e-03, -1.98412698412549389e-04, 2.75573192191536246e-06,   // [4..11] S0..S7
    -2.50521076157725835e-08, 1.60589772331473175e-10, -7.64396776493426846e-13, 2.73141320376900007e-15};
#define kInvPi kTrig[0]
#define kPiA kTrig[1]
#define kPiB kTrig[2]
#define kPiC kTrig[3]
#define kS0 kTrig[4]
#define kS1 kTrig[5]
#define kS2 kTrig[6]
#define kS3 kTrig[7]
#define kS4 kTrig[8]
#define kS5 kTrig[9]
#define kS6 kTrig[10]
#define kS7 kTrig[11]

// |x| outside the fast range of the trig kernels: >= 2^20, Inf, NaN
__device__ __forceinline__ bool trig_slow(double x) {
    return ((uint32_t)__double2hiint(x) & 0x7fffffffu) >= 0x41300000u;
}

// (-1)^k sin(r) for |r| <= pi/2 (a little beyond is fine); k in the low bit of `k`
__device__ __forceinline__ double sin_kernel(double r, uint32_t k) {
    // sin is odd: fold the parity into the argument
    const int rhi = __double2hiint(r) ^ (int)(k << 31);
    r = __hiloint2double(rhi, __double2loint(r));
    const double z = __dmul_rn(r, r);
    double p = kS7;
    p = __fma_rn(p, z, kS6);
    p = __fma_rn(p, z, kS5);
    p = __fma_rn(p, z, kS4);
    p = __fma_rn(p, z, kS3);
    p = __fma_rn(p, z, kS2);
    p = __fma_rn(p, z, kS1);
    p = __fma_rn(p, z, kS0);
    const double v = __fma_rn(__dmul_rn(r, z), p, r);
    // the result carries the sign of r; only differs from v for r = -0, where the fma yields +0
    return __hiloint2double(__double2hiint(v) | (rhi & (int)0x80000000), __double2loint(v));
}

__device__ __forceinline__ double sin_fast(double x) {
    const double t = __fma_rn(x, kInvPi, kMagic);
    const double kd = __dsub_rn(t, kMagic);
    double r = __fma_rn(kd, -kPiA, x);
    r = __fma_rn(kd, -kPiB, r);
    r = __fma_rn(kd, -kPiC, r);
    return sin_kernel(r, (uint32_t)__double2loint(t));
}
__device__ __forceinline__ double cos_fast(double x) {
    const double t = __dadd_rn(__fma_rn(x, kInvPi, -0.5), kMagic);
    const double hd = __dadd_rn(__dsub_rn(t, kMagic), 0.5);
    double r = __fma_rn(hd, -kPiA, x);
    r = __fma_rn(hd, -kPiB, r);
    r = __fma_rn(hd, -kPiC, r);
    return sin_kernel(r, (uint32_t)__double2loint(t) + 1u);
}

// ---- P-wide entry points -------------------------------------------------------------------------
// The fast path runs branch-free over all P points of a thread (the compiler interleaves the P
// dependent FMA chains: instruction-level parallelism instead of occupancy); points outside the fast
// range are then recomputed one by one with the CUDA math library.  Every point's result depends only
// on its own input -- never on which other points share the thread.
template <int P>
__device__ __forceinline__ void sin_n(const double (&x)[P], double (&out)[P]) {
    bool slow = false;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        out[p] = sin_fast(x[p]);
        slow |= trig_slow(x[p]);
    }
    if (slow) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (trig_slow(x[p])) out[p] = ::sin(x[p]);
    }
}
template <int P>
__device__ __forceinline__ void cos_n(const double (&x)[P], double (&out)[P]) {
    bool slow = false;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        out[p] = cos_fast(x[p]);
        slow |= trig_slow(x[p]);
    }
    if (slow) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (trig_slow(x[p])) out[p] = ::cos(x[p]);
    }
}
// sin AND cos of one argument with ONE reduction: k = rint(x * 2/pi), r = x - k pi/2 (three-term Cody-Waite, parts
// positive so that -0 survives), |r| <= pi/4; sin r = r + r z S(z), cos r = 1 + z (z C(z) - 1/2), z = r^2, S and C of
// degree 5 (tools/fit_leanmath.py: |S error| z < 1.3e-17, |C error| z^2 < 5e-19); the quadrant k mod 4 swaps and
// negates the two.  20 FP64 instructions for both values (sin_fast + cos_fast: 32) plus 2 selects and 3 logic
// operations per value.  <= 1.51 ulp against the exact value for |x| < 2^20 (tools/leanmath_check.c), i.e. slightly
// MORE accurate than sin_fast / cos_fast -- and therefore not bit-identical with them: a program whose Sin(x) and
// Cos(x) the lowering merges into one SinCos(x) differs from the unmerged one by at most 2 ulp per value.
__device__ __constant__ double kQuad[16] = {
    6.36619772367581382e-01,  // [0] 2/pi
    1.57079632679489656e+00,  // [1] pi/2, part A
    6.12323399573676480e-17,  // [2] pi/2, part B (rounded down)
    1.08285667392191403e-32,  // [3] pi/2, part C
    -1.66666666666666657e-01, 8.33333333333094450e-03, -1.98412698367513636e-04, 2.75573160988126868e-06,   // [4..9] S
    -2.50511310657360491e-08, 1.59180731629230736e-10,
    4.16666666666666644e-02, -1.38888888888873932e-03, 2.48015872987611761e-05, -2.75573172693902722e-07,   // [10..15] C
    2.08761457809035442e-09, -1.13825973098618771e-11};
__device__ __forceinline__ void sincos_fast(double x, double &sn, double &cs) {
    const double t = __fma_rn(x, kQuad[0], kMagic);
    const uint32_t q = (uint32_t)__double2loint(t);
    const double kd = __dsub_rn(t, kMagic);
    double r = __fma_rn(kd, -kQuad[1], x);
    r = __fma_rn(kd, -kQuad[2], r);
    r = __fma_rn(kd, -kQuad[3], r);
    const double z = __dmul_rn(r, r);
    double ps = kQuad[9], pc = kQuad[15];
    ps = __fma_rn(ps, z, kQuad[8]);   pc = __fma_rn(pc, z, kQuad[14]);
    ps = __fma_rn(ps, z, kQuad[7]);   pc = __fma_rn(pc, z, kQuad[13]);
    ps = __fma_rn(ps, z, kQuad[6]);   pc = __fma_rn(pc, z, kQuad[12]);
    ps = __fma_rn(ps, z, kQuad[5]);   pc = __fma_rn(pc, z, kQuad[11]);
    ps = __fma_rn(ps, z, kQuad[4]);   pc = __fma_rn(pc, z, kQuad[10]);
    const double s0 = __fma_rn(__dmul_rn(r, z), ps, r);
    const double c0 = __fma_rn(z, __fma_rn(z, pc, -0.5), 1.0);
    // sin r carries the sign of r also for r = -0 (where the fma yields +0)
    const int s_hi = (__double2hiint(s0) & 0x7fffffff) | (__double2hiint(r) & (int)0x80000000);
    const bool odd = (q & 1u) != 0u;
    const int a_hi = odd ? __double2hiint(c0) : s_hi, a_lo = odd ? __double2loint(c0) : __double2loint(s0);
    const int b_hi = odd ? s_hi : __double2hiint(c0), b_lo = odd ? __double2loint(s0) : __double2loint(c0);
    sn = __hiloint2double(a_hi ^ (int)((q & 2u) << 30), a_lo);
    cs = __hiloint2double(b_hi ^ (int)(((q + 1u) & 2u) << 30), b_lo);
}

template <int P>
__device__ __forceinline__ void sincos_n(const double (&x)[P], double (&s)[P], double (&c)[P]) {
    bool slow = false;
#pragma unroll
    for (int p = 0; p < P; ++p) {
#ifdef SAF_SINCOS_SEPARATE // build knob: the round-1 pair of separate kernels (32 FP64 instructions, bit-identical to Sin / Cos)
        s[p] = sin_fast(x[p]);
        c[p] = cos_fast(x[p]);
#else
        sincos_fast(x[p], s[p], c[p]);
#endif
        slow |= trig_slow(x[p]);
    }
    if (slow) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (trig_slow(x[p])) ::sincos(x[p], &s[p], &c[p]);
    }
}

// exp
__device__ __constant__ double kExp[14] = {
    1.44269504088896338700e+00,  // [0] log2(e)
    6.93147180369123816490e-01,  // [1] ln2 hi
    1.90821492927058770002e-10,  // [2] ln2 lo
    5.00000000000000000e-01, 1.66666666666666713e-01, 4.16666666666666713e-02, 8.33333333332612891e-03,   // [3..13] E0..E10
    1.38888888888837438e-03, 1.98412698748407683e-04, 2.48015873255621253e-05, 2.75572553746587440e-06,
    2.75572736248664783e-07, 2.51052276365517599e-08, 2.09146945603515922e-09};
#define kLog2e kExp[0]
#define kLn2Hi kExp[1]
#define kLn2Lo kExp[2]
#define kE0 kExp[3]
#define kE1 kExp[4]
#define kE2 kExp[5]
#define kE3 kExp[6]
#define kE4 kExp[7]
#define kE5 kExp[8]
#define kE6 kExp[9]
#define kE7 kExp[10]
#define kE8 kExp[11]
#define kE9 kExp[12]
#define kE10 kExp[13]

__device__ __forceinline__ double exp_fast(double x) {
    const double t = __fma_rn(x, kLog2e, kMagic);
    const double kd = __dsub_rn(t, kMagic);
    double r = __fma_rn(kd, -kLn2Hi, x);
    r = __fma_rn(kd, -kLn2Lo, r);
    double p = kE10;
    p = __fma_rn(p, r, kE9);
    p = __fma_rn(p, r, kE8);
    p = __fma_rn(p, r, kE7);
    p = __fma_rn(p, r, kE6);
    p = __fma_rn(p, r, kE5);
    p = __fma_rn(p, r, kE4);
    p = __fma_rn(p, r, kE3);
    p = __fma_rn(p, r, kE2);
    p = __fma_rn(p, r, kE1);
    p = __fma_rn(p, r, kE0);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    // scale by 2^k through the exponent field (results stay normal for |x| < 700)
    return __hiloint2double(__double2hiint(p) + (__double2loint(t) << 20), __double2loint(p));
}
__device__ __forceinline__ bool exp_slow(double x) {
    return ((uint32_t)__double2hiint(x) & 0x7fffffffu) >= 0x4085e000u; // |x| >= 700, Inf, NaN
}
template <int P>
__device__ __forceinline__ void exp_n(const double (&x)[P], double (&out)[P]) {
    bool slow = false;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        out[p] = exp_fast(x[p]);
        slow |= exp_slow(x[p]);
    }
    if (slow) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (exp_slow(x[p])) out[p] = ::exp(x[p]);
    }
}

} // namespace lm
} // namespace saf
