// kernel.cu -- the sm_100a interpreter kernel for SymbAnaFis device programs.
//
// One thread evaluates P points (P = 1, 2 or 4), 128 threads per block.  Every lane of every warp runs
// the SAME program on different points, so the program counter, the fetched micro-instruction and the
// opcode branch are warp-uniform: there is no divergence in the dispatch, only inside the data-dependent
// slow paths of a few special functions.
//
//   program image               : packed micro-instructions + constants (uop.h / pack.cpp), staged once per
//                                 block into SHARED memory (one 16-byte broadcast load fetches opcode,
//                                 destination and two operands); programs whose image does not fit beside
//                                 the register file are read from global memory through the read-only path
//   live values ("registers")   : shared memory laid out [slot][pair][thread] as f64x2 (f64 for P=1):
//                                 consecutive lanes touch consecutive 16-byte words -> conflict free;
//                                 operands arrive as byte offsets: one integer add per access
//   accumulator                 : the result of the previous instruction stays in real registers; the
//                                 lowering (lower.cpp) routes single-use values through it so they
//                                 never touch shared memory
//   dispatch                    : one dense switch (ptxas emits BRX jump tables for it) over micro-ops that
//                                 are specialised by operand kind (S slot / K constant / A accumulator):
//                                 handlers contain no operand decoding; P points per thread amortise it
//   hot / cold split            : the dispatch loop (hot_run) is a function of its own that contains only
//                                 the arithmetic micro-ops and the lean transcendentals -- ~70 registers, no
//                                 spills; pow, powi, the 61 builtins and the rare operand-kind combinations
//                                 run in cold_op, memory to memory, between two entries of hot_run.  Keeping
//                                 them out of the loop's function keeps ptxas' register allocation of the
//                                 loop independent of their needs (with them inline it used 128 registers,
//                                 spilled the program counter and copied ACC on every dispatch).
//   hot transcendentals         : issue-lean sin/cos/sincos/exp (leanmath.cuh: FP64 instructions plus
//                                 three integer instructions per point)
//   columns                     : struct-of-arrays f64, coalesced 128-bit loads/stores (64-bit for
//                                 P=1), streaming cache hints
//
// Arithmetic handlers use the explicit-rounding intrinsics (__dadd_rn, __dmul_rn, __fma_rn, ...) so
// that nvcc can never contract a Mul followed by an Add: the reference distinguishes Mul;Add from
// MulAdd (macros.rs:69-75 vs :128-138).  -fmad stays at its default because the CUDA math library
// needs it (see devmath.cuh).
#include "kernel.cuh"

#include <cstdio>

#include "devmath.cuh"
#include "leanmath.cuh"

// resident blocks per SM the register allocation aims for, per points-per-thread variant (tuning knobs)
#ifndef SAF_MINB_P4
#define SAF_MINB_P4 6
#endif
#ifndef SAF_MINB_P2
#define SAF_MINB_P2 6
#endif
#ifndef SAF_MINB_P1
#define SAF_MINB_P1 6
#endif

namespace saf {

namespace {

template <int P>
struct Pt {
    double v[P];
};

__device__ __forceinline__ double ld_stream(const double *p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void ld_stream2(const double *p, double &a, double &b) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}
__device__ __forceinline__ void st_stream(double *p, double v) {
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_stream2(double *p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

extern __shared__ __align__(16) unsigned char saf_smem[];

// Shared-memory register file.  For P >= 2 a thread's points are stored as P/2 pairs; pair g of the slot
// at byte offset `off` lives at  off + g * PAIR_BYTES + tid * 16  (`my` already includes tid * 16).
template <int P>
struct Rf {
    static constexpr uint32_t PAIR_BYTES = BLOCK_THREADS * 16;
    __device__ __forceinline__ static char *mine() {
        return reinterpret_cast<char *>(saf_smem) + threadIdx.x * (P == 1 ? 8 : 16);
    }
    __device__ __forceinline__ static void load(const char *my, uint32_t off, Pt<P> &o) {
        const char *p = my + off;
        if constexpr (P == 1) {
            o.v[0] = *reinterpret_cast<const double *>(p);
        } else {
#pragma unroll
            for (int g = 0; g < P / 2; ++g) {
                double2 t = *reinterpret_cast<const double2 *>(p + g * PAIR_BYTES);
                o.v[2 * g] = t.x;
                o.v[2 * g + 1] = t.y;
            }
        }
    }
    __device__ __forceinline__ static void store(char *my, uint32_t off, const Pt<P> &o) {
        char *p = my + off;
        if constexpr (P == 1) {
            *reinterpret_cast<double *>(p) = o.v[0];
        } else {
#pragma unroll
            for (int g = 0; g < P / 2; ++g)
                *reinterpret_cast<double2 *>(p + g * PAIR_BYTES) = make_double2(o.v[2 * g], o.v[2 * g + 1]);
        }
    }
};

#define SAF_FOR_P _Pragma("unroll") for (int p = 0; p < P; ++p)

__device__ __forceinline__ uint2 split64(unsigned long long w) { return make_uint2((uint32_t)w, (uint32_t)(w >> 32)); }

// Per-block context behind the register file in shared memory: what hot_run needs besides the image.
struct BlockCtx {
    uint32_t n_params;
    uint32_t result_k_mask;
    int32_t acc_off;
    uint32_t code_off;
    int32_t param_off[MAX_COLS];
    int32_t result_off[MAX_OUTS];
};

} // namespace

// Points of a thread.  P == 1: i = tile_base + tid.  P >= 2: pair g covers points
// tile_base + g * 2 * BLOCK + 2 * tid + {0, 1}: every 128-bit access of a warp is fully coalesced.
template <int P>
__device__ __forceinline__ unsigned long long point_index(unsigned long long tile_base, int p) {
    if constexpr (P == 1) return tile_base + threadIdx.x;
    else return tile_base + (unsigned long long)(p >> 1) * (2 * BLOCK_THREADS) + 2 * threadIdx.x + (p & 1);
}

// ---- cold micro-ops ---------------------------------------------------------------------------------
// pow, powi, 1/expm1, the builtins (61 FnOps of devmath.cuh), operand staging and the rare operand-kind
// combinations: decoded at run time, out of line, memory to memory.  ACC is read from / written to the
// reserved slot `acc_off`; hot_run stores ACC there before it returns and reloads it when re-entered.
template <int P>
__device__ __noinline__ void cold_op(const unsigned char *img, uint32_t cur, int32_t acc_off, int32_t x0_off,
                                     int32_t x1_off) {
    using RF = Rf<P>;
    char *const my = RF::mine();
    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(img + cur);
    const uint2 q0 = split64(w[0]), q1 = split64(w[1]), q2 = split64(w[2]), q3 = split64(w[3]);
    const uint32_t uop = q0.x, fa = q1.x, fb = q1.y, fc = q2.x, imm = q2.y, kinds = q3.x;
    const int32_t fd = (int32_t)q0.y;
    auto konst = [&](uint32_t off) -> double { return *reinterpret_cast<const double *>(img + off); };
    auto any = [&](uint32_t kind, uint32_t f, Pt<P> &o) {
        if (kind == KIND_S) {
            RF::load(my, f, o);
        } else if (kind == KIND_K) {
            const double k = konst(f);
            SAF_FOR_P o.v[p] = k;
        } else {
            RF::load(my, (uint32_t)acc_off, o);
        }
    };
    Pt<P> a, b, c, r;
    switch (uop) {
    case U_G + G_POW:
        any(kinds & 3u, fa, a); any((kinds >> 2) & 3u, fb, b);
        SAF_FOR_P r.v[p] = pow(a.v[p], b.v[p]);
        break;
    case U_G + G_DIV:
        any(kinds & 3u, fa, a); any((kinds >> 2) & 3u, fb, b);
        SAF_FOR_P r.v[p] = __ddiv_rn(a.v[p], b.v[p]);
        break;
    case U_G + G_POWI:
        any(kinds & 3u, fa, a);
        SAF_FOR_P r.v[p] = dm::powi(a.v[p], (int)imm).v;
        break;
    case U_G + G_RECIPEXPM1:
        any(kinds & 3u, fa, a);
        if (fb != NO_PREFIX) {
            const double k1 = konst(fb), k2 = konst(fb + 8);
            SAF_FOR_P a.v[p] = __fma_rn(a.v[p], k1, k2);
        }
        SAF_FOR_P r.v[p] = __ddiv_rn(1.0, expm1(a.v[p]));
        break;
    case U_G + G_BUILTIN1:
        any(kinds & 3u, fa, a);
        SAF_FOR_P r.v[p] = dm::builtin1(imm, a.v[p]).v;
        break;
    case U_G + G_BUILTIN2:
        any(kinds & 3u, fa, a); any((kinds >> 2) & 3u, fb, b);
        SAF_FOR_P r.v[p] = dm::builtin2(imm, a.v[p], b.v[p]).v;
        break;
    case U_G + G_BUILTIN3: {
        any(kinds & 3u, fa, a);
        Pt<P> x0;
        RF::load(my, (uint32_t)x0_off, x0);
        RF::load(my, (uint32_t)x1_off, c);
        SAF_FOR_P r.v[p] = dm::builtin3(imm, x0.v[p], c.v[p], a.v[p]).v;
    } break;
    case U_G + G_BUILTIN4: {
        any(kinds & 3u, fa, a); any((kinds >> 2) & 3u, fb, b);
        Pt<P> x0;
        RF::load(my, (uint32_t)x0_off, x0);
        RF::load(my, (uint32_t)x1_off, c);
        SAF_FOR_P r.v[p] = dm::builtin4(imm, x0.v[p], c.v[p], a.v[p], b.v[p]).v;
    } break;
    case U_G + G_ARG2: // stages two operands for the next Builtin3/4: ACC untouched
        any(kinds & 3u, fa, a); any((kinds >> 2) & 3u, fb, b);
        RF::store(my, (uint32_t)x0_off, a);
        RF::store(my, (uint32_t)x1_off, b);
        return;
    case U_G + G_FMA: case U_G + G_FMS: case U_G + G_FNMA: case U_G + G_FNMS: {
        any(kinds & 3u, fa, a); any((kinds >> 2) & 3u, fb, b); any((kinds >> 4) & 3u, fc, c);
        const uint32_t t = uop - (U_G + G_FMA);
        SAF_FOR_P {
            const double x = (t & 2u) ? -a.v[p] : a.v[p];
            const double z = (t & 1u) ? -c.v[p] : c.v[p];
            r.v[p] = __fma_rn(x, b.v[p], z);
        }
    } break;
    default: SAF_FOR_P r.v[p] = dm::nan_(); break;
    }
    RF::store(my, (uint32_t)acc_off, r);
    if (fd >= 0) RF::store(my, (uint32_t)fd, r);
}

// ---- the dispatch loop ------------------------------------------------------------------------------
// Runs micro-instructions from `pc` -- over as many iterations of the program as remain, feeding results
// back into the parameters between iterations -- until the program ends for the last time (returns pc = 0)
// or a cold micro-op is met (returns its pc; ACC has been written to its reserved slot).  The iteration
// counter travels in the high word.  IMG_SMEM: the image sits in shared memory at byte offset `img_s`,
// otherwise in global memory at `img_g`.
template <int P, bool IMG_SMEM>
__device__ __noinline__ unsigned long long hot_run(uint32_t img_s, const unsigned char *img_g, uint32_t ctx_off,
                                                   uint32_t pc, uint32_t it, uint32_t iters) {
    using RF = Rf<P>;
    char *const my = RF::mine();
    const BlockCtx *const ctx = reinterpret_cast<const BlockCtx *>(saf_smem + ctx_off);
    const unsigned char *const img_sm = saf_smem + img_s;

    auto img128 = [&](uint32_t off) -> uint4 { // 16 bytes of the image (warp-uniform address)
        if constexpr (IMG_SMEM) return *reinterpret_cast<const uint4 *>(img_sm + off);
        else return __ldg(reinterpret_cast<const uint4 *>(img_g + off));
    };
    auto img32 = [&](uint32_t off) -> uint32_t {
        if constexpr (IMG_SMEM) return *reinterpret_cast<const uint32_t *>(img_sm + off);
        else return __ldg(reinterpret_cast<const uint32_t *>(img_g + off));
    };
    auto konst = [&](uint32_t off) -> double {
        if constexpr (IMG_SMEM) return *reinterpret_cast<const double *>(img_sm + off);
        else return __ldg(reinterpret_cast<const double *>(img_g + off));
    };

    Pt<P> acc;
    RF::load(my, (uint32_t)ctx->acc_off, acc);

    for (;;) {
        for (;;) {
            const uint4 q = img128(pc);
            const uint32_t uop = q.x;
            const int32_t fd = (int32_t)q.y;
            const uint32_t fa = q.z, fb = q.w;
            const uint32_t cur = pc;
            pc += UINSTR_BYTES;

            Pt<P> a, b, c;

#define LS(x, f) RF::load(my, (f), x);
#define LK(x, f) { const double k_ = konst(f); SAF_FOR_P x.v[p] = k_; }
#define LA(x) x = acc;
#define FC (img32(cur + 16)) /* third operand offset */

// light unary: a from a slot / from ACC
#define U1CASE(OP, EXPR)                                                                                   \
    case U_U1 + 2 * ((OP) - D_FIRST_LIGHT_UNARY): LS(a, fa) SAF_FOR_P { const double x = a.v[p]; acc.v[p] = (EXPR); } break; \
    case U_U1 + 2 * ((OP) - D_FIRST_LIGHT_UNARY) + 1: SAF_FOR_P { const double x = acc.v[p]; acc.v[p] = (EXPR); } break;
// heavy unary: operand from a slot, {plain, affine prefix}; BODY maps a.v -> acc.v and never reads ACC
#define H1PREFIX { const double k1_ = konst(fb), k2_ = konst(fb + 8); SAF_FOR_P a.v[p] = __fma_rn(a.v[p], k1_, k2_); }
#define H1CASE(H, BODY)                                                                                    \
    case U_H1 + 2 * (H) + 0: LS(a, fa) { BODY } break;                                                     \
    case U_H1 + 2 * (H) + 1: LS(a, fa) H1PREFIX { BODY } break;
// commutative binary: SS SK SA KA AA
#define BCCASE(I, EXPR)                                                                                    \
    case U_BC + 5 * (I) + BC_SS: LS(a, fa) LS(b, fb) SAF_FOR_P { const double x = a.v[p], y = b.v[p]; acc.v[p] = (EXPR); } break; \
    case U_BC + 5 * (I) + BC_SK: LS(a, fa) { const double y = konst(fb); SAF_FOR_P { const double x = a.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_BC + 5 * (I) + BC_SA: LS(a, fa) SAF_FOR_P { const double x = a.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } break; \
    case U_BC + 5 * (I) + BC_KA: { const double x = konst(fa); SAF_FOR_P { const double y = acc.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_BC + 5 * (I) + BC_AA: SAF_FOR_P { const double x = acc.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } break;
// ordered binary: SS SK KS SA AS KA AK AA
#define BNCASE(I, EXPR)                                                                                    \
    case U_BN + 8 * (I) + BN_SS: LS(a, fa) LS(b, fb) SAF_FOR_P { const double x = a.v[p], y = b.v[p]; acc.v[p] = (EXPR); } break; \
    case U_BN + 8 * (I) + BN_SK: LS(a, fa) { const double y = konst(fb); SAF_FOR_P { const double x = a.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_BN + 8 * (I) + BN_KS: LS(b, fb) { const double x = konst(fa); SAF_FOR_P { const double y = b.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_BN + 8 * (I) + BN_SA: LS(a, fa) SAF_FOR_P { const double x = a.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } break; \
    case U_BN + 8 * (I) + BN_AS: LS(b, fb) SAF_FOR_P { const double x = acc.v[p], y = b.v[p]; acc.v[p] = (EXPR); } break; \
    case U_BN + 8 * (I) + BN_KA: { const double x = konst(fa); SAF_FOR_P { const double y = acc.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_BN + 8 * (I) + BN_AK: { const double y = konst(fb); SAF_FOR_P { const double x = acc.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_BN + 8 * (I) + BN_AA: SAF_FOR_P { const double x = acc.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } break;
// fused multiply-add family: x * y (+/-) z
#define TCASE(T, EXPR)                                                                                     \
    case U_T + 10 * (T) + T_SSS: { const uint32_t fc = FC; LS(a, fa) LS(b, fb) LS(c, fc) SAF_FOR_P { const double x = a.v[p], y = b.v[p], z = c.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_T + 10 * (T) + T_SSK: { const uint32_t fc = FC; LS(a, fa) LS(b, fb) const double z = konst(fc); SAF_FOR_P { const double x = a.v[p], y = b.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_T + 10 * (T) + T_SSA: { LS(a, fa) LS(b, fb) SAF_FOR_P { const double x = a.v[p], y = b.v[p], z = acc.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_T + 10 * (T) + T_SKS: { const uint32_t fc = FC; LS(a, fa) LS(c, fc) const double y = konst(fb); SAF_FOR_P { const double x = a.v[p], z = c.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_T + 10 * (T) + T_SKK: { const uint32_t fc = FC; LS(a, fa) const double y = konst(fb), z = konst(fc); SAF_FOR_P { const double x = a.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_T + 10 * (T) + T_SKA: { LS(a, fa) const double y = konst(fb); SAF_FOR_P { const double x = a.v[p], z = acc.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_T + 10 * (T) + T_SAS: { const uint32_t fc = FC; LS(a, fa) LS(c, fc) SAF_FOR_P { const double x = a.v[p], y = acc.v[p], z = c.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_T + 10 * (T) + T_SAK: { const uint32_t fc = FC; LS(a, fa) const double z = konst(fc); SAF_FOR_P { const double x = a.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_T + 10 * (T) + T_KAS: { const uint32_t fc = FC; LS(c, fc) const double x = konst(fa); SAF_FOR_P { const double y = acc.v[p], z = c.v[p]; acc.v[p] = (EXPR); } } break; \
    case U_T + 10 * (T) + T_KAK: { const uint32_t fc = FC; const double x = konst(fa), z = konst(fc); SAF_FOR_P { const double y = acc.v[p]; acc.v[p] = (EXPR); } } break;

            switch (uop) {
            case U_END: goto program_end;
            case U_LOADK: { const double k_ = konst(fa); SAF_FOR_P acc.v[p] = k_; } break;

            U1CASE(D_COPY, x)
            U1CASE(D_NEG, -x)
            U1CASE(D_SQUARE, __dmul_rn(x, x))
            U1CASE(D_CUBE, __dmul_rn(__dmul_rn(x, x), x))
            U1CASE(D_POW4, __dmul_rn(__dmul_rn(x, x), __dmul_rn(x, x)))
            U1CASE(D_POW3_2, __dmul_rn(x, __dsqrt_rn(x)))
            U1CASE(D_INVPOW3_2, __ddiv_rn(1.0, __dmul_rn(x, __dsqrt_rn(x))))
            U1CASE(D_INVSQRT, __ddiv_rn(1.0, __dsqrt_rn(x)))
            U1CASE(D_INVSQUARE, __ddiv_rn(1.0, __dmul_rn(x, x)))
            U1CASE(D_INVCUBE, __ddiv_rn(1.0, __dmul_rn(__dmul_rn(x, x), x)))
            U1CASE(D_RECIP, __ddiv_rn(1.0, x))
            U1CASE(D_SQRT, __dsqrt_rn(x))

            H1CASE(H1_SIN, lm::sin_n<P>(a.v, acc.v);)
            H1CASE(H1_COS, lm::cos_n<P>(a.v, acc.v);)
            H1CASE(H1_EXP, lm::exp_n<P>(a.v, acc.v);)
            H1CASE(H1_LN, SAF_FOR_P acc.v[p] = log(a.v[p]);)
            H1CASE(H1_EXPSQR, SAF_FOR_P a.v[p] = __dmul_rn(a.v[p], a.v[p]); lm::exp_n<P>(a.v, acc.v);)
            H1CASE(H1_EXPSQRNEG, SAF_FOR_P a.v[p] = -__dmul_rn(a.v[p], a.v[p]); lm::exp_n<P>(a.v, acc.v);)
            H1CASE(H1_EXPNEG, SAF_FOR_P a.v[p] = -a.v[p]; lm::exp_n<P>(a.v, acc.v);)
            H1CASE(H1_SINCOS, {
                const int32_t fc = (int32_t)FC;
                lm::sincos_n<P>(a.v, acc.v, c.v);
                if (fc >= 0) RF::store(my, (uint32_t)fc, c);
            })

            BCCASE(BC_ADD, __dadd_rn(x, y))
            BCCASE(BC_MUL, __dmul_rn(x, y))
            BCCASE(BC_NEGMUL, -__dmul_rn(x, y))
            BNCASE(BN_SUB, __dsub_rn(x, y))
            BNCASE(BN_DIV, __ddiv_rn(x, y))

            TCASE(T_FMA, __fma_rn(x, y, z))
            TCASE(T_FMS, __fma_rn(x, y, -z))
            TCASE(T_FNMA, __fma_rn(-x, y, z))
            TCASE(T_FNMS, __fma_rn(-x, y, -z))

            // cold micro-ops leave the loop: ACC goes to its slot, the caller runs cold_op and re-enters
            default:
                RF::store(my, (uint32_t)ctx->acc_off, acc);
                return ((unsigned long long)it << 32) | cur;
            }
            if (fd >= 0) RF::store(my, (uint32_t)fd, acc);
        }
    program_end:
        if (++it >= iters) return 0ull;
        // ---- in-kernel feedback: parameter j <- result j (result slots never alias parameter slots,
        // lower.cpp); the state stays on chip between iterations ----
        for (uint32_t j = 0; j < ctx->n_params; ++j) {
            const int32_t s = ctx->param_off[j];
            if (s < 0) continue;
            Pt<P> v;
            if ((ctx->result_k_mask >> j) & 1u) {
                const double k = konst((uint32_t)ctx->result_off[j]);
                SAF_FOR_P v.v[p] = k;
            } else {
                RF::load(my, (uint32_t)ctx->result_off[j], v);
            }
            RF::store(my, (uint32_t)s, v);
        }
        pc = ctx->code_off;
    }
}

// ---- the kernel ------------------------------------------------------------------------------------------
template <int P, bool IMG_SMEM>
__global__ void __launch_bounds__(BLOCK_THREADS, (P == 4 ? SAF_MINB_P4 : P == 2 ? SAF_MINB_P2 : SAF_MINB_P1))
saf_interp_kernel(const __grid_constant__ LaunchDesc d) {
    using RF = Rf<P>;
    char *const my = RF::mine();

    // ---- stage the block context (and the image) behind the register file ----
    const uint32_t ctx_off = d.n_slots * (uint32_t)(BLOCK_THREADS * P * 8);
    const uint32_t img_s = ctx_off + (uint32_t)sizeof(BlockCtx);
    {
        BlockCtx *ctx = reinterpret_cast<BlockCtx *>(saf_smem + ctx_off);
        if (threadIdx.x == 0) {
            ctx->n_params = d.n_params;
            ctx->result_k_mask = d.result_k_mask;
            ctx->acc_off = d.acc_off;
            ctx->code_off = d.code_off;
        }
        if (threadIdx.x < MAX_COLS) ctx->param_off[threadIdx.x] = d.param_off[threadIdx.x];
        if (threadIdx.x < MAX_OUTS) ctx->result_off[threadIdx.x] = d.result_off[threadIdx.x];
        if constexpr (IMG_SMEM) {
            const uint4 *src = reinterpret_cast<const uint4 *>(d.image_global);
            uint4 *dst = reinterpret_cast<uint4 *>(saf_smem + img_s);
            for (uint32_t i = threadIdx.x; i < d.image_bytes / 16u; i += BLOCK_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();
    }
    const unsigned char *const img_generic = IMG_SMEM ? (saf_smem + img_s) : d.image_global;
    const uint32_t iters = (uint32_t)d.iters;

    constexpr unsigned long long TILE = (unsigned long long)BLOCK_THREADS * P;
    const unsigned long long n_tiles = (d.n_points + TILE - 1) / TILE;

    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long tile_base = tile * TILE;

        // ---- prologue: columns -> parameter slots (scalar.rs:182-207 rules) ----
        for (uint32_t j = 0; j < d.n_params; ++j) {
            const int32_t s = d.param_off[j];
            if (s < 0) continue;
            const double *col = d.cols[j];
            const unsigned long long len = d.col_len[j];
            Pt<P> v;
            SAF_FOR_P v.v[p] = 0.0; // missing column reads as 0.0
            if (col != nullptr && len != 0ull) {
                if constexpr (P == 1) {
                    const unsigned long long i = point_index<P>(tile_base, 0);
                    if (i < d.n_points) v.v[0] = ld_stream(col + (i < len ? i : len - 1));
                } else {
#pragma unroll
                    for (int g = 0; g < P / 2; ++g) {
                        const unsigned long long i = point_index<P>(tile_base, 2 * g);
                        if (i + 1 < len && i + 1 < d.n_points) {
                            ld_stream2(col + i, v.v[2 * g], v.v[2 * g + 1]);
                        } else {
                            if (i < d.n_points) v.v[2 * g] = ld_stream(col + (i < len ? i : len - 1));
                            if (i + 1 < d.n_points) v.v[2 * g + 1] = ld_stream(col + (i + 1 < len ? i + 1 : len - 1));
                        }
                    }
                }
            }
            RF::store(my, (uint32_t)s, v);
        }

        // ---- the program: hot_run until it ends, cold micro-ops in between ----
        {
            uint32_t pc = d.code_off, it = 0;
            for (;;) {
                const unsigned long long st = hot_run<P, IMG_SMEM>(img_s, d.image_global, ctx_off, pc, it, iters);
                pc = (uint32_t)st;
                if (pc == 0u) break;
                it = (uint32_t)(st >> 32);
                cold_op<P>(img_generic, pc, d.acc_off, d.x0_off, d.x1_off);
                pc += UINSTR_BYTES;
            }
        }

        // ---- epilogue: result slots -> output columns ----
        for (uint32_t k = 0; k < d.n_results; ++k) {
            Pt<P> v;
            if ((d.result_k_mask >> k) & 1u) {
                const double kv = *reinterpret_cast<const double *>(img_generic + (uint32_t)d.result_off[k]);
                SAF_FOR_P v.v[p] = kv;
            } else {
                RF::load(my, (uint32_t)d.result_off[k], v);
            }
            double *out = d.outs[k];
            if constexpr (P == 1) {
                const unsigned long long i = point_index<P>(tile_base, 0);
                if (i < d.n_points) st_stream(out + i, v.v[0]);
            } else {
#pragma unroll
                for (int g = 0; g < P / 2; ++g) {
                    const unsigned long long i = point_index<P>(tile_base, 2 * g);
                    if (i + 1 < d.n_points) {
                        st_stream2(out + i, v.v[2 * g], v.v[2 * g + 1]);
                    } else if (i < d.n_points) {
                        st_stream(out + i, v.v[2 * g]);
                    }
                }
            }
        }
    }
}

// ---- FP64 pipe peak probe ---------------------------------------------------------------------------
// The roofline of this path is the FP64 pipe (BASELINE.json north_star), whose peak is not in
// MEASURED_PEAKS.json.  This kernel measures it on the box: 8 independent DFMA chains per thread,
// enough warps to saturate the pipe; bench.py divides lane-instructions by the CUDA-event time.
__global__ void __launch_bounds__(256) saf_fp64_peak_kernel(double *sink, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[0] = s; // never true: keeps the chains alive
}

cudaError_t launch_fp64_peak(double *sink, int blocks, int iters, cudaStream_t stream) {
    saf_fp64_peak_kernel<<<blocks, 256, 0, stream>>>(sink, iters, 0.5);
    return cudaGetLastError();
}

// ---- host side -------------------------------------------------------------------------------------

template <int P, bool IMG_SMEM>
static cudaError_t launch_one(const LaunchDesc &desc, const LaunchShape &shape, cudaStream_t stream) {
    auto kern = saf_interp_kernel<P, IMG_SMEM>;
    // opt in once per (instantiation, device) to the full dynamic shared-memory carve-out
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        int optin = 0;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    kern<<<shape.grid, shape.block, shape.smem_bytes, stream>>>(desc);
    return cudaGetLastError();
}

cudaError_t launch_interpreter(const LaunchDesc &desc, const LaunchShape &shape, cudaStream_t stream) {
#ifdef SAF_FAST_BUILD // developer builds: one instantiation only (tools/sass_probe.sh)
    if (shape.points_per_thread != 4 || !shape.image_in_smem) return cudaErrorNotSupported;
    return launch_one<4, true>(desc, shape, stream);
#else
    const bool s = shape.image_in_smem;
    switch (shape.points_per_thread) {
    case 4: return s ? launch_one<4, true>(desc, shape, stream) : launch_one<4, false>(desc, shape, stream);
    case 2: return s ? launch_one<2, true>(desc, shape, stream) : launch_one<2, false>(desc, shape, stream);
    default: return s ? launch_one<1, true>(desc, shape, stream) : launch_one<1, false>(desc, shape, stream);
    }
#endif
}

// register file + block context (+ the image when it is staged in shared memory)
size_t interpreter_smem_bytes(uint32_t n_slots, int points_per_thread, size_t image_bytes_in_smem) {
    return (size_t)n_slots * (size_t)BLOCK_THREADS * (size_t)points_per_thread * 8 + sizeof(BlockCtx) +
           ((image_bytes_in_smem + 15) & ~(size_t)15);
}

static int resident_blocks_for(int P) { return P == 4 ? SAF_MINB_P4 : P == 2 ? SAF_MINB_P2 : SAF_MINB_P1; }

cudaError_t choose_points_per_thread(int device, uint32_t n_slots, unsigned long long n_points, bool aligned16,
                                     int force_points_per_thread, int *points_per_thread) {
    int sms = 0, smem_optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    const size_t smem_sm = (size_t)smem_optin; // per-block opt-in limit ~= per-SM capacity minus 1 KB
    // More points per thread amortise the dispatch; they need (a) 16-byte aligned columns, (b) enough
    // points to still give every SM several blocks, (c) a register file that leaves >= 2 blocks per SM.
    // n_slots + 3: room for the reserved slots the packer adds.
    auto fits = [&](int p, int min_blocks) {
        return (interpreter_smem_bytes(n_slots + 3, p, 0) + 1024) * (size_t)min_blocks <= smem_sm + 1024;
    };
    auto enough = [&](int p) {
        return n_points >= (unsigned long long)sms * 3ull * (unsigned long long)BLOCK_THREADS * (unsigned long long)p;
    };
    int P = 1;
    if (aligned16 && enough(4) && fits(4, 2)) P = 4;
    else if (aligned16 && enough(2) && fits(2, 2)) P = 2;
    if (force_points_per_thread == 1 || (aligned16 && (force_points_per_thread == 2 || force_points_per_thread == 4)))
        P = force_points_per_thread;
    while (P > 1 && !fits(P, 1)) P >>= 1;
    if (!fits(P, 1)) return cudaErrorInvalidConfiguration;
    *points_per_thread = P;
    return cudaSuccess;
}

cudaError_t choose_shape(int device, uint32_t n_slots, size_t image_bytes, unsigned long long n_points, int P,
                         LaunchShape *shape) {
    int sms = 0, smem_optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    const size_t rf = interpreter_smem_bytes(n_slots, P, 0);
    if (rf > (size_t)smem_optin) return cudaErrorInvalidConfiguration;
    // the image rides in shared memory when it is small next to the register file (it must not cost a
    // resident block) and always when it is tiny
    bool in_smem = image_bytes <= IMAGE_SMEM_ALWAYS;
    if (!in_smem && image_bytes <= IMAGE_SMEM_MAX) {
        const size_t with = interpreter_smem_bytes(n_slots, P, image_bytes);
        const size_t blocks_without = ((size_t)smem_optin + 1024) / (rf + 1024);
        const size_t blocks_with = with <= (size_t)smem_optin ? ((size_t)smem_optin + 1024) / (with + 1024) : 0;
        const size_t want = (size_t)resident_blocks_for(P);
        in_smem = blocks_with >= (blocks_without < want ? blocks_without : want);
    }
    size_t smem = interpreter_smem_bytes(n_slots, P, in_smem ? image_bytes : 0);
    if (smem > (size_t)smem_optin) {
        in_smem = false;
        smem = rf;
    }
    size_t resident = (size_t)resident_blocks_for(P); // __launch_bounds__ register budget
    const size_t by_smem = ((size_t)smem_optin + 1024) / (smem + 1024);
    if (by_smem < resident) resident = by_smem;
    if (resident < 1) resident = 1;
    const unsigned long long tile = (unsigned long long)BLOCK_THREADS * P;
    const unsigned long long n_tiles = (n_points + tile - 1) / tile;
    unsigned long long grid = (unsigned long long)sms * resident;
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    shape->points_per_thread = P;
    shape->block = BLOCK_THREADS;
    shape->grid = (int)grid;
    shape->smem_bytes = smem;
    shape->image_in_smem = in_smem;
    return cudaSuccess;
}

} // namespace saf
