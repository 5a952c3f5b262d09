// kernel_r.cuh -- dispatch loop of the REGISTER machine (ruop.h); included by kernel.cu inside namespace saf.
//
// Live values sit in eight named registers R0..R7 of four points each (64 of the thread's 32-bit registers) that
// stay allocated across every handler of the loop; the register numbers of an operation are part of its opcode,
// so a handler is nothing but its FP64 instructions: `Rd = Rd op X` with X another register, a constant (one
// broadcast shared-memory load) or a spilled value (two 16-byte shared-memory loads).  Heavy operations (the
// transcendentals, division, square root) work in place on R0 -- their code is too long to replicate per register.
// The registers do NOT survive a return from hot_run_r: pack_r.cpp stores every live register to its slot in front
// of a cold micro-op.
#pragma once

#include "ruop.h"

namespace {

#define R_FOR_D(M, ...) M(0, __VA_ARGS__) M(1, __VA_ARGS__) M(2, __VA_ARGS__) M(3, __VA_ARGS__) M(4, __VA_ARGS__) M(5, __VA_ARGS__) M(6, __VA_ARGS__) M(7, __VA_ARGS__)
#define R_FOR_X(M, ...) M(0, __VA_ARGS__) M(1, __VA_ARGS__) M(2, __VA_ARGS__) M(3, __VA_ARGS__) M(4, __VA_ARGS__) M(5, __VA_ARGS__) M(6, __VA_ARGS__) M(7, __VA_ARGS__)
#define R_FOR_PAIR(M, ...)                                                                                                                   \
    M(0, 0, __VA_ARGS__) M(0, 1, __VA_ARGS__) M(0, 2, __VA_ARGS__) M(0, 3, __VA_ARGS__) M(0, 4, __VA_ARGS__) M(0, 5, __VA_ARGS__) M(0, 6, __VA_ARGS__) M(0, 7, __VA_ARGS__) \
    M(1, 1, __VA_ARGS__) M(1, 2, __VA_ARGS__) M(1, 3, __VA_ARGS__) M(1, 4, __VA_ARGS__) M(1, 5, __VA_ARGS__) M(1, 6, __VA_ARGS__) M(1, 7, __VA_ARGS__) \
    M(2, 2, __VA_ARGS__) M(2, 3, __VA_ARGS__) M(2, 4, __VA_ARGS__) M(2, 5, __VA_ARGS__) M(2, 6, __VA_ARGS__) M(2, 7, __VA_ARGS__)                      \
    M(3, 3, __VA_ARGS__) M(3, 4, __VA_ARGS__) M(3, 5, __VA_ARGS__) M(3, 6, __VA_ARGS__) M(3, 7, __VA_ARGS__)                                           \
    M(4, 4, __VA_ARGS__) M(4, 5, __VA_ARGS__) M(4, 6, __VA_ARGS__) M(4, 7, __VA_ARGS__)                                                                \
    M(5, 5, __VA_ARGS__) M(5, 6, __VA_ARGS__) M(5, 7, __VA_ARGS__)                                                                                     \
    M(6, 6, __VA_ARGS__) M(6, 7, __VA_ARGS__)                                                                                                          \
    M(7, 7, __VA_ARGS__)

// one handler: case label, patcher marker, body, back to the dispatch
#define R_H(V, ...) case (V): R_ENTER(V) { __VA_ARGS__ } R_NEXT(V)

// operand X: register x / constant at byte offset fa / slot at byte offset fa
#define R_X_K(NAME, F) Pt<P> NAME; { const double k_ = konst(F); SAF_FOR_P NAME.v[p] = k_; }
#define R_X_S(NAME, F) Pt<P> NAME; RF::load(my, (F), NAME);

// ---- Rd = X ----
#define R_MOV_DX(x, d) R_H(RU_MOV + (d) * 10 + (x), r##d = r##x;)
#define R_MOV_D(d, _)                                                                                      \
    R_FOR_X(R_MOV_DX, d)                                                                                   \
    R_H(RU_MOV + (d) * 10 + XK_K, { const double k_ = konst(fa); SAF_FOR_P r##d.v[p] = k_; })              \
    R_H(RU_MOV + (d) * 10 + XK_S, RF::load(my, fa, r##d);)
// ---- S[a] = Rd ----
#define R_ST_D(d, _) R_H(RU_ST + (d), RF::store(my, fa, r##d);)
// ---- Rd = Rd op X ----
#define RB_E0(D, X) __dadd_rn(D, X)
#define RB_E1(D, X) __dmul_rn(D, X)
#define RB_E2(D, X) __dsub_rn(D, X)
#define RB_E3(D, X) __dsub_rn(X, D)
#define RB_E4(D, X) (-__dmul_rn(D, X))
#define R_BIN_DX(x, d, op, E) R_H(RU_BIN + ((op) * 8 + (d)) * 10 + (x), SAF_FOR_P r##d.v[p] = E(r##d.v[p], r##x.v[p]);)
#define R_BIN_D(d, op, E)                                                                                  \
    R_FOR_X(R_BIN_DX, d, op, E)                                                                            \
    R_H(RU_BIN + ((op) * 8 + (d)) * 10 + XK_K, { const double k_ = konst(fa); SAF_FOR_P r##d.v[p] = E(r##d.v[p], k_); }) \
    R_H(RU_BIN + ((op) * 8 + (d)) * 10 + XK_S, R_X_S(x_, fa) SAF_FOR_P r##d.v[p] = E(r##d.v[p], x_.v[p]);)
// ---- Rd = f(Rd) ----
#define R_UN_D(d, u, EXPR) R_H(RU_UN + (u) * 8 + (d), SAF_FOR_P { const double x = r##d.v[p]; r##d.v[p] = (EXPR); })
// ---- fused multiply-add family: x * y (+/-) z ----
#define RF_E0(X, Y, Z) __fma_rn(X, Y, Z)
#define RF_E1(X, Y, Z) __fma_rn(X, Y, -(Z))
#define RF_E2(X, Y, Z) __fma_rn(-(X), Y, Z)
#define RF_E3(X, Y, Z) __fma_rn(-(X), Y, -(Z))
// (on R0 only, see ruop.h)
#define R_FRR_P(a, b, s, E) R_H(RU_FRR + (s) * 36 + r_pair(a, b), SAF_FOR_P r0.v[p] = E(r##a.v[p], r##b.v[p], r0.v[p]);)
#define R_FRR_S(s, E) R_FOR_PAIR(R_FRR_P, s, E)
#define R_FKA_X(x, s, E) R_H(RU_FKA + (s) * 8 + (x), { const double k_ = konst(fa); SAF_FOR_P r0.v[p] = E(r##x.v[p], k_, r0.v[p]); })
#define R_FKA_S(s, E) R_FOR_X(R_FKA_X, s, E)
#define R_FKM_X(x, s, E) R_H(RU_FKM + (s) * 10 + (x), { const double k_ = konst(fa); SAF_FOR_P r0.v[p] = E(r0.v[p], k_, r##x.v[p]); })
#define R_FKM_S(s, E)                                                                                      \
    R_FOR_X(R_FKM_X, s, E)                                                                                 \
    R_H(RU_FKM + (s) * 10 + XK_K, { const double k_ = konst(fa), z_ = konst(fb); SAF_FOR_P r0.v[p] = E(r0.v[p], k_, z_); }) \
    R_H(RU_FKM + (s) * 10 + XK_S, { const double k_ = konst(fa); R_X_S(z_, fb) SAF_FOR_P r0.v[p] = E(r0.v[p], k_, z_.v[p]); })
// ---- heavy operations, in place on R0 ----
// R1 = f(R0): argument and result in different registers -- the n-wide leanmath kernels re-read their argument on
// the slow path, and with an in-place form ptxas splits R0's live range at the loop head (8 moves per dispatch)
#define R_H0PREFIX { const double k1_ = konst(fb), k2_ = konst(fb + 8); SAF_FOR_P a_.v[p] = __fma_rn(a_.v[p], k1_, k2_); }
#define R_H0CASE(H, ...)                                                                                   \
    R_H(RU_H0 + 2 * (H) + 0, const Pt<P> &a_ = r0; __VA_ARGS__)                                            \
    R_H(RU_H0 + 2 * (H) + 1, Pt<P> a_ = r0; R_H0PREFIX __VA_ARGS__)
#define R_U0CASE(U, EXPR) R_H(RU_U0 + (U), SAF_FOR_P { const double x = r0.v[p]; r0.v[p] = (EXPR); })
#define R_DIV_X(x, _) R_H(RU_DIV + (x), SAF_FOR_P r0.v[p] = __ddiv_rn(r0.v[p], r##x.v[p]);) \
                      R_H(RU_RDIV + (x), SAF_FOR_P r0.v[p] = __ddiv_rn(r##x.v[p], r0.v[p]);)

} // namespace

// Runs register-machine micro-instructions from code offset `pc_code`; same contract as hot_run: returns HOT_DONE at
// RU_EXIT, or the code offset of a cold micro-op (low word) and the iteration counter (high word).
template <bool K_SMEM, bool THREADED>
__device__ __noinline__ unsigned long long hot_run_r(uint32_t smem32, uint32_t ctx_off, const unsigned char *img_g,
                                                     uint32_t pc_code, uint32_t it, uint32_t iters) {
    constexpr int P = R_POINTS;
    using RF = Rf<P>;
    const uint32_t my = RF::mine(smem32);
    const uint32_t ctx = smem32 + ctx_off;
    const uint32_t k_sm = ctx + (uint32_t)sizeof(BlockCtx);
    const uint32_t code_off = lds_u32(ctx + (uint32_t)offsetof(BlockCtx, code_off));
    const uint32_t code_bytes = lds_u32(ctx + (uint32_t)offsetof(BlockCtx, code_bytes));
    const uint32_t win_sm = k_sm + (K_SMEM ? code_off : 0u);
    const unsigned char *const code_g = img_g + code_off;
    uint32_t win_g = pc_code & ~(CODE_WINDOW_BYTES - 1u);
    uint32_t pc = win_sm + (pc_code & (CODE_WINDOW_BYTES - 1u));

    auto konst = [&](uint32_t off) -> double {
        if constexpr (K_SMEM) {
            double k;
            lds64(k_sm + off, k);
            return k;
        } else {
            return __ldg(reinterpret_cast<const double *>(img_g + off));
        }
    };

    Pt<P> r0, r1, r2, r3, r4, r5, r6, r7;
    SAF_FOR_P { r0.v[p] = 0.0; r1.v[p] = 0.0; r2.v[p] = 0.0; r3.v[p] = 0.0; r4.v[p] = 0.0; r5.v[p] = 0.0; r6.v[p] = 0.0; r7.v[p] = 0.0; }

    // Threaded dispatch (see hot_run): a handler fetches the NEXT micro-instruction as its first action, so that the
    // fetch overlaps its own work, and ends with its own jump through the table.  The opcode travels one instruction
    // ahead of its operands (word 0 holds the opcode of the following instruction).
    uint32_t uop = lds_u32(pc + 28);
    uint32_t fa, fb, uop_now;
    uint4 qn;
#define R_ROTATE { fa = qn.z; fb = qn.w; pc += UINSTR_BYTES; uop_now = uop; uop = qn.x; }
#define R_ENTER(V) SAF_CASE(V) if constexpr (THREADED) { qn = lds_u128(pc); }
#define R_NEXT(id) { if constexpr (THREADED) { R_ROTATE SAF_REDISPATCH(uop_now, id) continue; } else { break; } }
#define R_RESTART(id) { uop = lds_u32(pc + 28); if constexpr (THREADED) { qn = lds_u128(pc); R_ROTATE SAF_REDISPATCH(uop_now, id) } continue; }
    if constexpr (THREADED) { qn = lds_u128(pc); R_ROTATE }
    for (;;) {
        if constexpr (!THREADED) { qn = lds_u128(pc); R_ROTATE }
        SAF_DISPATCH_N(uop_now, RU_COUNT)
        switch (uop_now) {
        case RU_END: SAF_CASE(RU_END) // a = code offset of the loop body; the last iteration falls through
            if (++it < iters) {
                const uint32_t w = fa & ~(CODE_WINDOW_BYTES - 1u);
                if (w != win_g) {
                    win_g = w;
                    stage_window(win_sm, code_g, win_g, code_bytes);
                }
                pc = win_sm + (fa & (CODE_WINDOW_BYTES - 1u));
            }
            R_RESTART(RU_END)
        case RU_EXIT: SAF_CASE(RU_EXIT) return HOT_DONE;
        case RU_NEXTWIN: SAF_CASE(RU_NEXTWIN)
            win_g += CODE_WINDOW_BYTES;
            stage_window(win_sm, code_g, win_g, code_bytes);
            pc = win_sm;
            R_RESTART(RU_NEXTWIN)

        R_FOR_D(R_MOV_D, 0)
        R_FOR_D(R_ST_D, 0)
        R_FOR_D(R_BIN_D, RB_ADD, RB_E0)
        R_FOR_D(R_BIN_D, RB_MUL, RB_E1)
        R_FOR_D(R_BIN_D, RB_SUB, RB_E2)
        R_FOR_D(R_BIN_D, RB_RSUB, RB_E3)
        R_FOR_D(R_BIN_D, RB_NEGMUL, RB_E4)
        R_FOR_D(R_UN_D, RUN_NEG, -x)
        R_FOR_D(R_UN_D, RUN_CUBE, __dmul_rn(__dmul_rn(x, x), x))
        R_FOR_D(R_UN_D, RUN_POW4, __dmul_rn(__dmul_rn(x, x), __dmul_rn(x, x)))
        R_FRR_S(T_FMA, RF_E0)
        R_FRR_S(T_FMS, RF_E1)
        R_FRR_S(T_FNMA, RF_E2)
        R_FRR_S(T_FNMS, RF_E3)
        R_FKA_S(T_FMA, RF_E0)
        R_FKA_S(T_FMS, RF_E1)
        R_FKA_S(T_FNMA, RF_E2)
        R_FKA_S(T_FNMS, RF_E3)
        R_FKM_S(T_FMA, RF_E0)
        R_FKM_S(T_FMS, RF_E1)
        R_FKM_S(T_FNMA, RF_E2)
        R_FKM_S(T_FNMS, RF_E3)

        R_H0CASE(H1_SIN, lm::sin_n<P>(a_.v, r1.v);)
        R_H0CASE(H1_COS, lm::cos_n<P>(a_.v, r1.v);)
        R_H0CASE(H1_EXP, lm::exp_n<P>(a_.v, r1.v);)
        R_H0CASE(H1_LN, SAF_FOR_P r1.v[p] = log(a_.v[p]);)
        R_H0CASE(H1_EXPSQR, Pt<P> b_; SAF_FOR_P b_.v[p] = __dmul_rn(a_.v[p], a_.v[p]); lm::exp_n<P>(b_.v, r1.v);)
        R_H0CASE(H1_EXPSQRNEG, Pt<P> b_; SAF_FOR_P b_.v[p] = -__dmul_rn(a_.v[p], a_.v[p]); lm::exp_n<P>(b_.v, r1.v);)
        R_H0CASE(H1_EXPNEG, Pt<P> b_; SAF_FOR_P b_.v[p] = -a_.v[p]; lm::exp_n<P>(b_.v, r1.v);)
        R_H0CASE(H1_SINCOS, lm::sincos_n<P>(a_.v, r1.v, r2.v);)

        R_U0CASE(RU0_POW3_2, __dmul_rn(x, __dsqrt_rn(x)))
        R_U0CASE(RU0_INVPOW3_2, __ddiv_rn(1.0, __dmul_rn(x, __dsqrt_rn(x))))
        R_U0CASE(RU0_INVSQRT, __ddiv_rn(1.0, __dsqrt_rn(x)))
        R_U0CASE(RU0_INVSQUARE, __ddiv_rn(1.0, __dmul_rn(x, x)))
        R_U0CASE(RU0_INVCUBE, __ddiv_rn(1.0, __dmul_rn(__dmul_rn(x, x), x)))
        R_U0CASE(RU0_RECIP, __ddiv_rn(1.0, x))
        R_U0CASE(RU0_SQRT, __dsqrt_rn(x))
        R_FOR_X(R_DIV_X, 0)
        R_H(RU_DIV + XK_K, { const double k_ = konst(fa); SAF_FOR_P r0.v[p] = __ddiv_rn(r0.v[p], k_); })
        R_H(RU_DIV + XK_S, R_X_S(x_, fa) SAF_FOR_P r0.v[p] = __ddiv_rn(r0.v[p], x_.v[p]);)
        R_H(RU_RDIV + XK_K, { const double k_ = konst(fa); SAF_FOR_P r0.v[p] = __ddiv_rn(k_, r0.v[p]); })
        R_H(RU_RDIV + XK_S, R_X_S(x_, fa) SAF_FOR_P r0.v[p] = __ddiv_rn(x_.v[p], r0.v[p]);)

        // cold micro-ops leave the loop (the packer has stored every live register)
        default: SAF_CASE_DEFAULT
            return ((unsigned long long)it << 32) | (win_g + (pc - UINSTR_BYTES - win_sm));
        }
    }
#undef R_NEXT
#undef R_ENTER
}
