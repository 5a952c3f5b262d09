// spec_prelude.cuh -- fixed part of a SPECIALISED kernel (spec.cpp): column tiles in, output tiles out.
//
// A specialised kernel is the device program of ONE saf_program (saf_isa.h) written out as straight-line sm_100a code:
// every slot a register, every constant an immediate, no fetch, no dispatch, no shared-memory register file.  The
// statements are the interpreter's handlers verbatim -- the same explicit-rounding intrinsics in the same order, the
// same leanmath.cuh / devmath.cuh leaf functions compiled from the same text -- so both machines return the same bits
// (tests/test_gpu_specialized.py).  What is left to write by hand is what surrounds the statements: this file.
//
// Point ownership is the interpreter's (kernel.cu point_index): thread t of a tile owns, for pair g, the points
// tile_base + g * 2 * BLOCK + 2 * t + {0, 1}; every 128-bit access of a warp is fully coalesced.  Column rules are
// those of classic_tile (scalar.rs:182-207): a missing column reads 0.0, a short column repeats its last element.
//
// Compiled at run time by NVRTC (jit.cpp) with SPEC_P (points per thread: 1, 2, 4), SPEC_BLOCK and SPEC_MINB defined
// by the generated text that includes it; nvcc compiles it too (`make spec_check`) so that errors show up at build time.
#pragma once
#include <cstdint>

#include "devmath.cuh"
#include "leanmath.cuh"

namespace saf {
namespace sp {

constexpr int MAX_COLS_ = 32;
constexpr int MAX_OUTS_ = 32;

// kernel parameter block of every specialised kernel (host mirror: spec.h SpecDesc)
struct SpecDesc {
    unsigned long long n_points;
    unsigned long long iters;
    const double *cols[MAX_COLS_];
    unsigned long long col_len[MAX_COLS_];
    double *outs[MAX_OUTS_];
};

#define SAF_P _Pragma("unroll") for (int p = 0; p < SPEC_P; ++p)

// Plain (coherent) loads: saf_iterate_device and in-place evaluation pass the same buffers as columns and outputs.
#ifdef SAF_HOST_SHIM // tests/spec_host: the generated text compiled by g++ for the CPU test-suite (never in the product)
inline double ld1(const double *p) { return *p; }
inline void ld2(const double *p, double &a, double &b) { a = p[0]; b = p[1]; }
inline void st1(double *p, double v) { *p = v; }
inline void st2(double *p, double a, double b) { p[0] = a; p[1] = b; }
#else
__device__ __forceinline__ double ld1(const double *p) {
    double r;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void ld2(const double *p, double &a, double &b) {
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st1(double *p, double v) { asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void st2(double *p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
#endif

template <int P, int BLOCK>
__device__ __forceinline__ unsigned long long point_index(unsigned long long tile_base, int p) {
    if constexpr (P == 1) return tile_base + threadIdx.x;
    else return tile_base + (unsigned long long)(p >> 1) * (2u * BLOCK) + 2u * threadIdx.x + (unsigned)(p & 1);
}

// column j of the tile -> v.  `full`: the whole tile lies inside the batch (the common case: no per-point tests)
template <int P, int BLOCK>
__device__ __forceinline__ void load_col(const SpecDesc &d, int j, unsigned long long tile_base, bool full, double (&v)[P]) {
    const double *col = d.cols[j];
    const unsigned long long len = d.col_len[j];
#pragma unroll
    for (int p = 0; p < P; ++p) v[p] = 0.0;
    if (col == nullptr || len == 0ull) return;
    if constexpr (P == 1) {
        const unsigned long long i = point_index<P, BLOCK>(tile_base, 0);
        if (i < d.n_points) v[0] = ld1(col + (i < len ? i : len - 1));
    } else {
        const bool vec = full && len >= d.n_points && (((unsigned long long)col & 15ull) == 0ull);
#pragma unroll
        for (int g = 0; g < P / 2; ++g) {
            const unsigned long long i = point_index<P, BLOCK>(tile_base, 2 * g);
            if (vec) {
                ld2(col + i, v[2 * g], v[2 * g + 1]);
            } else {
                if (i < d.n_points) v[2 * g] = ld1(col + (i < len ? i : len - 1));
                if (i + 1 < d.n_points) v[2 * g + 1] = ld1(col + (i + 1 < len ? i + 1 : len - 1));
            }
        }
    }
}

template <int P, int BLOCK>
__device__ __forceinline__ void store_out(const SpecDesc &d, int k, unsigned long long tile_base, bool full, const double (&v)[P]) {
    double *out = d.outs[k];
    if constexpr (P == 1) {
        const unsigned long long i = point_index<P, BLOCK>(tile_base, 0);
        if (i < d.n_points) st1(out + i, v[0]);
    } else {
        const bool vec = full && (((unsigned long long)out & 15ull) == 0ull);
#pragma unroll
        for (int g = 0; g < P / 2; ++g) {
            const unsigned long long i = point_index<P, BLOCK>(tile_base, 2 * g);
            if (vec) {
                st2(out + i, v[2 * g], v[2 * g + 1]);
            } else {
                if (i < d.n_points) st1(out + i, v[2 * g]);
                if (i + 1 < d.n_points) st1(out + i + 1, v[2 * g + 1]);
            }
        }
    }
}

// Out-of-line leaves (SPEC_OUTLINE, set by the generator for programs with many transcendentals): ONE copy of each lean
// kernel per module instead of one per call site -- a 110-term gradient inlines 330 of them otherwise (minutes of
// compile time, megabytes of code that no instruction cache holds).  Operands travel by value: registers, not memory.
template <int P>
struct Vec {
    double v[P];
};
template <int P>
struct Vec2 {
    double s[P], c[P];
};
#ifdef SPEC_OUTLINE
#define SAF_LEAF __device__ __noinline__
#else
#define SAF_LEAF __device__ __forceinline__
#endif
SAF_LEAF Vec<SPEC_P> sin_v(Vec<SPEC_P> a) { Vec<SPEC_P> r; lm::sin_n<SPEC_P>(a.v, r.v); return r; }
SAF_LEAF Vec<SPEC_P> cos_v(Vec<SPEC_P> a) { Vec<SPEC_P> r; lm::cos_n<SPEC_P>(a.v, r.v); return r; }
SAF_LEAF Vec<SPEC_P> exp_v(Vec<SPEC_P> a) { Vec<SPEC_P> r; lm::exp_n<SPEC_P>(a.v, r.v); return r; }
SAF_LEAF Vec2<SPEC_P> sincos_v(Vec<SPEC_P> a) { Vec2<SPEC_P> r; lm::sincos_n<SPEC_P>(a.v, r.s, r.c); return r; }
// speculative leaves (spec.cpp `speculate`): the fast path only; `slow` collects the largest |argument| high word seen,
// which the generated code compares once per iteration with the fast-range limits of leanmath.cuh (trig_slow, exp_slow)
__device__ __forceinline__ unsigned hi_abs(double x) { return (unsigned)__double2hiint(x) & 0x7fffffffu; }
__device__ __forceinline__ Vec<SPEC_P> sin_f(Vec<SPEC_P> a, unsigned &slow) {
    Vec<SPEC_P> r;
#pragma unroll
    for (int p = 0; p < SPEC_P; ++p) { r.v[p] = lm::sin_fast(a.v[p]); slow = max(slow, hi_abs(a.v[p])); }
    return r;
}
__device__ __forceinline__ Vec<SPEC_P> cos_f(Vec<SPEC_P> a, unsigned &slow) {
    Vec<SPEC_P> r;
#pragma unroll
    for (int p = 0; p < SPEC_P; ++p) { r.v[p] = lm::cos_fast(a.v[p]); slow = max(slow, hi_abs(a.v[p])); }
    return r;
}
__device__ __forceinline__ Vec<SPEC_P> exp_f(Vec<SPEC_P> a, unsigned &slow) {
    Vec<SPEC_P> r;
#pragma unroll
    for (int p = 0; p < SPEC_P; ++p) { r.v[p] = lm::exp_fast(a.v[p]); slow = max(slow, hi_abs(a.v[p])); }
    return r;
}
__device__ __forceinline__ Vec2<SPEC_P> sincos_f(Vec<SPEC_P> a, unsigned &slow) {
    Vec2<SPEC_P> r;
#pragma unroll
    for (int p = 0; p < SPEC_P; ++p) {
#ifdef SAF_SINCOS_SEPARATE
        r.s[p] = lm::sin_fast(a.v[p]);
        r.c[p] = lm::cos_fast(a.v[p]);
#else
        lm::sincos_fast(a.v[p], r.s[p], r.c[p]);
#endif
        slow = max(slow, hi_abs(a.v[p]));
    }
    return r;
}
static_assert(true, "fast-range limits used by the generated code: trig 0x41300000 (2^20), exp 0x4085e000 (700)");
#ifdef SPEC_OUTLINE
SAF_LEAF Vec2<2 * SPEC_P> sincos_v2(Vec<2 * SPEC_P> a) { Vec2<2 * SPEC_P> r; lm::sincos_n<2 * SPEC_P>(a.v, r.s, r.c); return r; }
#endif
SAF_LEAF double log_(double a) { return log(a); }
SAF_LEAF double div_(double a, double b) { return __ddiv_rn(a, b); }

// the cold leaves, called exactly as kernel.cu's cold_op calls them
__device__ __forceinline__ double pow_(double a, double b) { return pow(a, b); }
__device__ __forceinline__ double powi_(double a, int n) { return dm::powi(a, n).v; }
__device__ __forceinline__ double recip_expm1_(double a) { return __ddiv_rn(1.0, expm1(a)); }
__device__ __forceinline__ double builtin1_(uint32_t fn, double a) { return dm::builtin1(fn, a).v; }
__device__ __forceinline__ double builtin2_(uint32_t fn, double a, double b) { return dm::builtin2(fn, a, b).v; }
__device__ __forceinline__ double builtin3_(uint32_t fn, double a, double b, double c) { return dm::builtin3(fn, a, b, c).v; }
__device__ __forceinline__ double builtin4_(uint32_t fn, double a, double b, double c, double e) { return dm::builtin4(fn, a, b, c, e).v; }

} // namespace sp
} // namespace saf
