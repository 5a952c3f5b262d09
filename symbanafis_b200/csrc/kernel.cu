// kernel.cu -- the sm_100a interpreter kernel for SymbAnaFis device programs.
//
// One thread evaluates P points (P = 1, 2 or 4), 128 threads per block.  Every lane of every warp runs
// the SAME device program on different points, so the program counter, the fetched instruction word
// and the opcode branch are warp-uniform: there is no divergence in the dispatch, only inside the
// data-dependent loops of a few special functions.
//
//   program words + constants : kernel parameter space = constant bank (ConstBlob), read through the
//                               constant cache with uniform addresses; global-memory fallback
//                               (read-only path) for programs that do not fit
//   live values ("registers")  : shared memory laid out [slot][pair][thread] as f64x2 (f64 for P=1):
//                               consecutive lanes touch consecutive 16-byte words -> conflict free;
//                               the slot stride is a compile-time constant (fixed block size)
//   accumulator                : the result of the previous instruction stays in real registers; the
//                               lowering (lower.cpp) routes single-use values through it so they
//                               never touch shared memory
//   dispatch                   : one dense switch over opcodes that are specialised by operand kind
//                               for the cheap operations (no per-operand branches in their handlers);
//                               P points per thread amortise it
//   hot transcendentals        : issue-lean sin/cos/sincos/exp (leanmath.cuh), everything else from
//                               the CUDA math library / devmath.cuh
//   columns                    : struct-of-arrays f64, coalesced 128-bit loads/stores (64-bit for
//                               P=1), streaming cache hints
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
#define SAF_MINB_P4 3
#endif
#ifndef SAF_MINB_P2
#define SAF_MINB_P2 4
#endif
#ifndef SAF_MINB_P1
#define SAF_MINB_P1 4
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

// Shared-memory register file.  For P >= 2 a thread's points are stored as P/2 pairs; pair g of slot i
// lives at  i * SLOT_BYTES + g * PAIR_BYTES + tid * 16.  A field is index << 2 | kind, so
// (field & ~3) * (SLOT_BYTES / 4) is the byte offset of an S field: one LOP3 + one IMAD.
template <int P>
struct Rf {
    static constexpr uint32_t SLOT_BYTES = BLOCK_THREADS * P * 8;
    static constexpr uint32_t PAIR_BYTES = BLOCK_THREADS * 16;
    __device__ __forceinline__ static void load(const char *my, uint32_t field, Pt<P> &o) {
        const char *p = my + (field & 0xfffcu) * (SLOT_BYTES / 4);
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
    __device__ __forceinline__ static void store(char *my, uint32_t field, const Pt<P> &o) {
        char *p = my + (field & 0xfffcu) * (SLOT_BYTES / 4);
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

} // namespace

// Points of a thread.  P == 1: i = tile_base + tid.  P >= 2: pair g covers points
// tile_base + g * 2 * BLOCK + 2 * tid + {0, 1}: every 128-bit access of a warp is fully coalesced.
template <int P>
__device__ __forceinline__ unsigned long long point_index(unsigned long long tile_base, int p) {
    if constexpr (P == 1) return tile_base + threadIdx.x;
    else return tile_base + (unsigned long long)(p >> 1) * (2 * BLOCK_THREADS) + 2 * threadIdx.x + (p & 1);
}

// CODE_IN_BLOB / K_IN_BLOB: program words / constants come from the constant bank (kernel
// parameter `blob`), otherwise from global memory through the read-only path.
template <int P, int WORDS, bool CODE_IN_BLOB, bool K_IN_BLOB>
__global__ void __launch_bounds__(BLOCK_THREADS, (P == 4 ? SAF_MINB_P4 : P == 2 ? SAF_MINB_P2 : SAF_MINB_P1))
saf_interp_kernel(const __grid_constant__ LaunchDesc d, const __grid_constant__ ConstBlob<WORDS> blob) {
    extern __shared__ __align__(16) unsigned char saf_smem[];
    char *const my = reinterpret_cast<char *>(saf_smem) + threadIdx.x * (P == 1 ? 8 : 16);
    using RF = Rf<P>;

    // trig coefficient table behind the register file
    double *const trig_tab = reinterpret_cast<double *>(saf_smem + (size_t)d.n_slots * RF::SLOT_BYTES);
    if (threadIdx.x < 14) trig_tab[threadIdx.x] = (&lm::kTrigTable[0][0])[threadIdx.x];
    __syncthreads();

    auto fetch_word = [&](uint32_t pc) -> uint2 {
        unsigned long long w;
        if constexpr (CODE_IN_BLOB) w = blob.w[pc];
        else w = __ldg(reinterpret_cast<const unsigned long long *>(d.code_global) + pc);
        return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
    };
    auto konst_at = [&](uint32_t idx) -> double { // constant table entry (warp-uniform load)
        if constexpr (K_IN_BLOB) return __longlong_as_double((long long)blob.w[d.n_code + idx]);
        else return __ldg(d.consts_global + idx);
    };
    auto konst = [&](uint32_t field) -> double { return konst_at(field >> KIND_BITS); };

    constexpr unsigned long long TILE = (unsigned long long)BLOCK_THREADS * P;
    const unsigned long long n_tiles = (d.n_points + TILE - 1) / TILE;

    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long tile_base = tile * TILE;

        // ---- prologue: columns -> parameter slots (scalar.rs:182-207 rules) ----
        for (uint32_t j = 0; j < d.n_params; ++j) {
            const uint32_t s = d.param_slot[j];
            if (s == 0xffffu) continue;
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
            RF::store(my, s << KIND_BITS, v);
        }

        Pt<P> acc, x0, x1;
        SAF_FOR_P { acc.v[p] = 0.0; x0.v[p] = 0.0; x1.v[p] = 0.0; }

        for (unsigned long long it = 0;;) {
            // ---- the interpreter loop: warp-uniform pc / opcode ----
            uint32_t pc = 0;
            for (;;) {
                const uint2 w = fetch_word(pc++);
                const uint32_t xop = w.x & 0xffu;
                const uint32_t fd = w.x >> 16;
                const uint32_t fa = w.y & 0xffffu;
                const uint32_t fb = w.y >> 16;

                // run-time operand fetch for the heavy opcodes
                auto any = [&](uint32_t f, Pt<P> &o) {
                    const uint32_t kind = f & KIND_MASK;
                    if (kind == KIND_S) {
                        RF::load(my, f, o);
                    } else if (kind == KIND_K) {
                        const double k = konst(f);
                        SAF_FOR_P o.v[p] = k;
                    } else {
                        o = acc;
                    }
                };
                // operand of a heavy unary opcode, with the optional affine prefix (saf_isa.h)
                auto unary_arg = [&](Pt<P> &o) {
                    any(fa, o);
                    const uint32_t mode = (w.x >> 8) & 0xffu;
                    if (mode != PREFIX_NONE) {
                        const uint32_t ki = fb >> KIND_BITS;
                        const double k1 = konst_at(ki);
                        if (mode == PREFIX_MUL) {
                            SAF_FOR_P o.v[p] = __dmul_rn(o.v[p], k1);
                        } else {
                            const double k2 = konst_at(ki + 1);
                            SAF_FOR_P o.v[p] = __fma_rn(o.v[p], k1, k2);
                        }
                    }
                };

                Pt<P> a, b, r;

#define SAF_A_S RF::load(my, fa, a);
#define SAF_A_K { const double k_ = konst(fa); SAF_FOR_P a.v[p] = k_; }
#define SAF_A_A a = acc;
#define SAF_B_S RF::load(my, fb, b);
#define SAF_B_K { const double k_ = konst(fb); SAF_FOR_P b.v[p] = k_; }
#define SAF_B_A b = acc;
// light unary: two forms (a from a slot / from ACC)
#define SAF_ULIGHT(OP, EXPR)                                                                       \
    case X_ULIGHT + 2 * ((OP) - D_FIRST_LIGHT_UNARY): SAF_A_S SAF_FOR_P { EXPR; } break;           \
    case X_ULIGHT + 2 * ((OP) - D_FIRST_LIGHT_UNARY) + 1: SAF_A_A SAF_FOR_P { EXPR; } break;
// light binary: eight forms (K,K is never emitted)
#define SAF_BCASE(OP, KA, KB, LA, LB, EXPR)                                                        \
    case X_BLIGHT + 9 * ((OP) - D_FIRST_LIGHT_BINARY) + 3 * (KA) + (KB): LA LB SAF_FOR_P { EXPR; } break;
#define SAF_BLIGHT(OP, EXPR)                                                                       \
    SAF_BCASE(OP, KIND_S, KIND_S, SAF_A_S, SAF_B_S, EXPR)                                          \
    SAF_BCASE(OP, KIND_S, KIND_K, SAF_A_S, SAF_B_K, EXPR)                                          \
    SAF_BCASE(OP, KIND_S, KIND_ACC, SAF_A_S, SAF_B_A, EXPR)                                        \
    SAF_BCASE(OP, KIND_K, KIND_S, SAF_A_K, SAF_B_S, EXPR)                                          \
    SAF_BCASE(OP, KIND_K, KIND_ACC, SAF_A_K, SAF_B_A, EXPR)                                        \
    SAF_BCASE(OP, KIND_ACC, KIND_S, SAF_A_A, SAF_B_S, EXPR)                                        \
    SAF_BCASE(OP, KIND_ACC, KIND_K, SAF_A_A, SAF_B_K, EXPR)                                        \
    SAF_BCASE(OP, KIND_ACC, KIND_ACC, SAF_A_A, SAF_B_A, EXPR)
#define SAF_UHEAVY(OP, EXPR)                                                                       \
    case X_UHEAVY + ((OP) - D_FIRST_HEAVY_UNARY): unary_arg(a); SAF_FOR_P { EXPR; } break;
#define SAF_BHEAVY(OP, EXPR)                                                                       \
    case X_BHEAVY + ((OP) - D_FIRST_HEAVY_BINARY): any(fa, a); any(fb, b); SAF_FOR_P { EXPR; } break;
#define SAF_TERN(OP, EXPR)                                                                         \
    case X_TERN + ((OP) - D_FIRST_TERNARY): {                                                      \
        const uint2 h = fetch_word(pc++);                                                          \
        Pt<P> c;                                                                                   \
        any(fa, a); any(fb, b); any(h.x & 0xffffu, c);                                             \
        SAF_FOR_P { EXPR; }                                                                        \
    } break;

                switch (xop) {
                case X_END: goto program_end;
                case X_LOADK: SAF_A_K r = a; break;

                SAF_ULIGHT(D_COPY, r.v[p] = a.v[p])
                SAF_ULIGHT(D_NEG, r.v[p] = -a.v[p])
                SAF_ULIGHT(D_SQUARE, r.v[p] = __dmul_rn(a.v[p], a.v[p]))
                SAF_ULIGHT(D_CUBE, r.v[p] = __dmul_rn(__dmul_rn(a.v[p], a.v[p]), a.v[p]))
                SAF_ULIGHT(D_POW4, double t_ = __dmul_rn(a.v[p], a.v[p]); r.v[p] = __dmul_rn(t_, t_))
                SAF_ULIGHT(D_POW3_2, r.v[p] = __dmul_rn(a.v[p], __dsqrt_rn(a.v[p])))
                SAF_ULIGHT(D_INVPOW3_2, r.v[p] = __ddiv_rn(1.0, __dmul_rn(a.v[p], __dsqrt_rn(a.v[p]))))
                SAF_ULIGHT(D_INVSQRT, r.v[p] = __ddiv_rn(1.0, __dsqrt_rn(a.v[p])))
                SAF_ULIGHT(D_INVSQUARE, r.v[p] = __ddiv_rn(1.0, __dmul_rn(a.v[p], a.v[p])))
                SAF_ULIGHT(D_INVCUBE, r.v[p] = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(a.v[p], a.v[p]), a.v[p])))
                SAF_ULIGHT(D_RECIP, r.v[p] = __ddiv_rn(1.0, a.v[p]))
                SAF_ULIGHT(D_SQRT, r.v[p] = __dsqrt_rn(a.v[p]))

                case X_UHEAVY + (D_SIN - D_FIRST_HEAVY_UNARY): unary_arg(a); lm::sin_n<P>(a.v, r.v, trig_tab); break;
                case X_UHEAVY + (D_COS - D_FIRST_HEAVY_UNARY): unary_arg(a); lm::cos_n<P>(a.v, r.v, trig_tab); break;
                case X_UHEAVY + (D_EXP - D_FIRST_HEAVY_UNARY): unary_arg(a); lm::exp_n<P>(a.v, r.v); break;
                SAF_UHEAVY(D_LN, r.v[p] = log(a.v[p]))
                SAF_UHEAVY(D_RECIPEXPM1, r.v[p] = __ddiv_rn(1.0, expm1(a.v[p])))
                case X_UHEAVY + (D_EXPSQR - D_FIRST_HEAVY_UNARY):
                    unary_arg(a);
                    SAF_FOR_P a.v[p] = __dmul_rn(a.v[p], a.v[p]);
                    lm::exp_n<P>(a.v, r.v);
                    break;
                case X_UHEAVY + (D_EXPSQRNEG - D_FIRST_HEAVY_UNARY):
                    unary_arg(a);
                    SAF_FOR_P a.v[p] = -__dmul_rn(a.v[p], a.v[p]);
                    lm::exp_n<P>(a.v, r.v);
                    break;
                case X_UHEAVY + (D_EXPNEG - D_FIRST_HEAVY_UNARY):
                    unary_arg(a);
                    SAF_FOR_P a.v[p] = -a.v[p];
                    lm::exp_n<P>(a.v, r.v);
                    break;
                case X_UHEAVY + (D_SINCOS - D_FIRST_HEAVY_UNARY): {
                    const uint2 h = fetch_word(pc++);
                    unary_arg(a);
                    Pt<P> cs;
                    lm::sincos_n<P>(a.v, r.v, cs.v, trig_tab);
                    if ((h.x & KIND_MASK) == KIND_S) RF::store(my, h.x & 0xffffu, cs);
                } break;
                case X_UHEAVY + (D_BUILTIN1 - D_FIRST_HEAVY_UNARY): {
                    const uint2 h = fetch_word(pc++);
                    any(fa, a);
                    SAF_FOR_P r.v[p] = dm::builtin1(h.y, a.v[p]).v;
                } break;
                case X_UHEAVY + (D_BUILTIN3 - D_FIRST_HEAVY_UNARY): {
                    const uint2 h = fetch_word(pc++);
                    any(fa, a);
                    SAF_FOR_P r.v[p] = dm::builtin3(h.y, x0.v[p], x1.v[p], a.v[p]).v;
                } break;

                SAF_BLIGHT(D_ADD, r.v[p] = __dadd_rn(a.v[p], b.v[p]))
                SAF_BLIGHT(D_SUB, r.v[p] = __dsub_rn(a.v[p], b.v[p]))
                SAF_BLIGHT(D_MUL, r.v[p] = __dmul_rn(a.v[p], b.v[p]))
                SAF_BLIGHT(D_NEGMUL, r.v[p] = -__dmul_rn(a.v[p], b.v[p]))

                SAF_BHEAVY(D_DIV, r.v[p] = __ddiv_rn(a.v[p], b.v[p]))
                SAF_BHEAVY(D_POW, r.v[p] = pow(a.v[p], b.v[p]))
                case X_BHEAVY + (D_POWI - D_FIRST_HEAVY_BINARY): {
                    const uint2 h = fetch_word(pc++);
                    any(fa, a);
                    const int n = (int)h.y;
                    SAF_FOR_P r.v[p] = dm::powi(a.v[p], n).v;
                } break;
                case X_BHEAVY + (D_BUILTIN2 - D_FIRST_HEAVY_BINARY): {
                    const uint2 h = fetch_word(pc++);
                    any(fa, a); any(fb, b);
                    SAF_FOR_P r.v[p] = dm::builtin2(h.y, a.v[p], b.v[p]).v;
                } break;
                case X_BHEAVY + (D_BUILTIN4 - D_FIRST_HEAVY_BINARY): {
                    const uint2 h = fetch_word(pc++);
                    any(fa, a); any(fb, b);
                    SAF_FOR_P r.v[p] = dm::builtin4(h.y, x0.v[p], x1.v[p], a.v[p], b.v[p]).v;
                } break;
                case X_BHEAVY + (D_ARG2 - D_FIRST_HEAVY_BINARY):
                    any(fa, x0);
                    any(fb, x1);
                    continue; // stages operands only: ACC and slots untouched

                SAF_TERN(D_FMA, r.v[p] = __fma_rn(a.v[p], b.v[p], c.v[p]))
                SAF_TERN(D_FMS, r.v[p] = __fma_rn(a.v[p], b.v[p], -c.v[p]))
                SAF_TERN(D_FNMA, r.v[p] = __fma_rn(-a.v[p], b.v[p], c.v[p]))
                SAF_TERN(D_FNMS, r.v[p] = __fma_rn(-a.v[p], b.v[p], -c.v[p]))
                default: SAF_FOR_P r.v[p] = dm::nan_(); break;
                }
                acc = r;
                if ((fd & KIND_MASK) == KIND_S) RF::store(my, fd, r);
            }
        program_end:

            if (++it >= d.iters) break;
            // ---- in-kernel feedback: parameter j <- result j (result slots never alias parameter
            // slots, lower.cpp), state stays on chip between iterations ----
            for (uint32_t j = 0; j < d.n_params; ++j) {
                const uint32_t s = d.param_slot[j];
                if (s == 0xffffu) continue;
                const uint32_t f = d.result_field[j];
                Pt<P> v;
                if ((f & KIND_MASK) == KIND_S) {
                    RF::load(my, f, v);
                } else {
                    const double k = konst(f);
                    SAF_FOR_P v.v[p] = k;
                }
                RF::store(my, s << KIND_BITS, v);
            }
        }

        // ---- epilogue: result slots -> output columns ----
        for (uint32_t k = 0; k < d.n_results; ++k) {
            const uint32_t f = d.result_field[k];
            Pt<P> v;
            if ((f & KIND_MASK) == KIND_S) {
                RF::load(my, f, v);
            } else {
                const double kv = konst(f);
                SAF_FOR_P v.v[p] = kv;
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

template <int P, int WORDS, bool CODE_IN_BLOB, bool K_IN_BLOB>
static cudaError_t launch_one(const LaunchDesc &desc, const uint64_t *words, size_t n_words,
                              const LaunchShape &shape, cudaStream_t stream) {
    auto kern = saf_interp_kernel<P, WORDS, CODE_IN_BLOB, K_IN_BLOB>;
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
    ConstBlob<WORDS> blob;
    if (CODE_IN_BLOB) {
        for (size_t i = 0; i < n_words; ++i) blob.w[i] = words[i];
    }
    kern<<<shape.grid, shape.block, shape.smem_bytes, stream>>>(desc, blob);
    return cudaGetLastError();
}

template <int P>
static cudaError_t launch_p(const LaunchDesc &desc, const uint64_t *words, size_t n_words, bool consts_in_blob,
                            const LaunchShape &shape, cudaStream_t stream) {
    if (words == nullptr || n_words > (size_t)BLOB_LARGE)
        return launch_one<P, 1, false, false>(desc, nullptr, 0, shape, stream);
    if (!consts_in_blob) return launch_one<P, BLOB_LARGE, true, false>(desc, words, n_words, shape, stream);
    if (n_words <= (size_t)BLOB_SMALL) return launch_one<P, BLOB_SMALL, true, true>(desc, words, n_words, shape, stream);
    return launch_one<P, BLOB_LARGE, true, true>(desc, words, n_words, shape, stream);
}

cudaError_t launch_interpreter(const LaunchDesc &desc, const uint64_t *blob_words, size_t n_blob_words,
                               bool consts_in_blob, const LaunchShape &shape, cudaStream_t stream) {
    switch (shape.points_per_thread) {
    case 4: return launch_p<4>(desc, blob_words, n_blob_words, consts_in_blob, shape, stream);
    case 2: return launch_p<2>(desc, blob_words, n_blob_words, consts_in_blob, shape, stream);
    default: return launch_p<1>(desc, blob_words, n_blob_words, consts_in_blob, shape, stream);
    }
}

size_t interpreter_smem_bytes(uint32_t n_slots, int points_per_thread) {
    return (size_t)n_slots * (size_t)BLOCK_THREADS * (size_t)points_per_thread * 8 + 128; // + trig table
}

cudaError_t choose_shape(int device, uint32_t n_slots, unsigned long long n_points, bool aligned16,
                         int force_points_per_thread, LaunchShape *shape) {
    int sms = 0, smem_optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    const size_t smem_sm = (size_t)smem_optin; // per-block opt-in limit ~= per-SM capacity minus 1 KB
    const int block = BLOCK_THREADS;

    // More points per thread amortise the dispatch; they need (a) 16-byte aligned columns, (b) enough
    // points to still give every SM several blocks, (c) a register file that leaves >= 2 blocks per SM.
    auto fits = [&](int p, int min_blocks) {
        return interpreter_smem_bytes(n_slots, p) * (size_t)min_blocks <= smem_sm;
    };
    auto enough = [&](int p) {
        return n_points >= (unsigned long long)sms * 3ull * (unsigned long long)block * (unsigned long long)p;
    };
    int P = 1;
    if (aligned16 && enough(4) && fits(4, 2)) P = 4;
    else if (aligned16 && enough(2) && fits(2, 2)) P = 2;
    if (force_points_per_thread == 1 || (aligned16 && (force_points_per_thread == 2 || force_points_per_thread == 4)))
        P = force_points_per_thread;
    if (!fits(P, 1)) {
        P = 1;
        if (!fits(1, 1)) return cudaErrorInvalidConfiguration;
    }

    const size_t smem = interpreter_smem_bytes(n_slots, P);
    size_t resident = (P == 4) ? SAF_MINB_P4 : (P == 2) ? SAF_MINB_P2 : SAF_MINB_P1; // __launch_bounds__ budget
    size_t by_smem = (smem_sm + 1024) / (smem + 1024);
    if (by_smem < resident) resident = by_smem;
    if (resident < 1) resident = 1;
    const unsigned long long tile = (unsigned long long)block * P;
    const unsigned long long n_tiles = (n_points + tile - 1) / tile;
    unsigned long long grid = (unsigned long long)sms * resident;
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    shape->points_per_thread = P;
    shape->block = block;
    shape->grid = (int)grid;
    shape->smem_bytes = smem;
    return cudaSuccess;
}

} // namespace saf
