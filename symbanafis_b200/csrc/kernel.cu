// kernel.cu -- the sm_100a interpreter kernel for SymbAnaFis device programs.
//
// One thread evaluates P points (P = 1, 2, 4 or 8), 128 threads per block (up to 512 in the wide two-points layout).  Every lane of every warp runs
// the SAME program on different points, so the program counter, the fetched micro-instruction and the
// opcode branch are warp-uniform: there is no divergence in the dispatch, only inside the data-dependent
// slow paths of a few special functions.
//
//   program image               : packed micro-instructions + constants (uop.h / pack.cpp).  Constants are
//                                 staged once per block into SHARED memory; the code runs out of a 16 KB
//                                 shared-memory WINDOW (one 16-byte broadcast load fetches next opcode,
//                                 destination and two operands) that is restaged from global memory at
//                                 U_NEXTWIN micro-ops -- programs of any length, no per-instruction bounds
//                                 checks, programs under 512 micro-instructions never restage
//   live values ("registers")   : shared memory laid out [slot][pair][thread] as f64x2 (f64 for P=1):
//                                 consecutive lanes touch consecutive 16-byte words -> conflict free;
//                                 operands arrive as byte offsets: one integer add per access
//   accumulator                 : the result of the previous instruction stays in real registers; the
//                                 lowering (lower.cpp) routes single-use values through it so they
//                                 never touch shared memory
//   dispatch                    : one jump table (brx.idx, written into the PTX by tools/patch_brx.py) over micro-ops
//                                 that are specialised by operand kind (S slot / K constant / A accumulator):
//                                 handlers contain no operand decoding; P points per thread amortise it.  Chains
//                                 of light operations are fused into single micro-ops by the packer (uop.h U_FM,
//                                 U_FQ, U_FS, U_FT): a dispatch costs more than the arithmetic it leads to
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
#include "ruop.h"

#include <cstddef>
#include <cstdio>

#include "devmath.cuh"
#include "leanmath.cuh"

namespace saf {

namespace {

template <int P>
struct Pt {
    double v[P];
};

// Plain (coherent) loads with a streaming hint, not ld.global.nc: saf_iterate_device and in-place evaluation pass
// the same buffers as columns and outputs, and PTX leaves a non-coherent read of an address the same kernel writes
// undefined (a broadcast column that aliases an output is re-read by later tiles).
__device__ __forceinline__ double ld_stream(const double *p) {
    double r;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void ld_stream2(const double *p, double &a, double &b) {
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_stream(double *p, double v) {
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_stream2(double *p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

extern __shared__ __align__(16) unsigned char saf_smem[];

// -DSAF_CHECKED (tools/build_variant.sh checked -DSAF_CHECKED): every shared-memory access of the interpreter -- they
// are raw 32-bit shared addresses in inline PTX, computed from offsets the packer wrote -- is checked against the
// block's dynamic shared-memory window and its natural alignment, every bulk copy against its column; a violation
// prints what and where and traps.  compute-sanitizer is closed on the GPU pool this was developed on; this build
// stands in for its memcheck on the accesses that the compiler cannot see (tools/gpu_checked.sh).
#ifdef SAF_CHECKED
__device__ __noinline__ void saf_check_fail(const char *what, uint32_t addr, uint32_t bytes, uint32_t base, uint32_t size) {
    printf("SAF_CHECKED: %s: shared address 0x%x (+%u bytes) outside [0x%x, 0x%x) or misaligned; block %d thread %d\n", what, addr,
           bytes, base, base + size, (int)blockIdx.x, (int)threadIdx.x);
    asm volatile("trap;");
}
__device__ __forceinline__ void saf_check(const char *what, uint32_t addr, uint32_t bytes) {
    uint32_t size;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(size));
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(saf_smem);
    if (addr < base || addr + bytes > base + size || (addr & (bytes - 1u)) != 0u) saf_check_fail(what, addr, bytes, base, size);
}
#define SAF_CHECK(what, addr, bytes) saf_check(what, addr, bytes)
#else
#define SAF_CHECK(what, addr, bytes)
#endif

// Shared-memory register file, addressed by 32-bit shared-space addresses.  For P >= 2 a thread's points are
// stored as P/2 pairs; pair g of the slot at byte offset `off` lives at  off + g * PAIR_BYTES + tid * 16
// (`my` = block's shared base + tid * 16).  The base arrives as an opaque function argument / is computed once:
// when the code mentions the shared-memory symbol itself ptxas re-derives the base (S2UR + UMOV + ULEA) on
// every access.
__device__ __forceinline__ void lds64(uint32_t addr, double &x) {
    SAF_CHECK("lds64", addr, 8u);
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr));
}
__device__ __forceinline__ void lds128(uint32_t addr, double &x, double &y) {
    SAF_CHECK("lds128", addr, 16u);
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void sts64(uint32_t addr, double x) {
    SAF_CHECK("sts64", addr, 8u);
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(x) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, double x, double y) {
    SAF_CHECK("sts128", addr, 16u);
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    SAF_CHECK("lds_u32", addr, 4u);
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    SAF_CHECK("lds_u128", addr, 16u);
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t smem_base32() { return (uint32_t)__cvta_generic_to_shared(saf_smem); }

// ---- bulk asynchronous copies (TMA, 1-D) + mbarrier: the streaming path of interp_body ------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    SAF_CHECK("bulk_load dst", dst_smem, 16u);
    SAF_CHECK("bulk_load end", dst_smem + bytes - 16u, 16u);
    SAF_CHECK("bulk_load mbarrier", bar, 8u);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes) {
    SAF_CHECK("bulk_store src", src_smem, 16u);
    SAF_CHECK("bulk_store end", src_smem + bytes - 16u, 16u);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// WIDE (one block of blockDim.x threads per SM): the stride between a thread's pairs is the block's width, a
// run-time value -- the accesses use [register + uniform register] addressing, no extra instruction.
template <int P, bool WIDE = false>
struct Rf {
    __device__ __forceinline__ static uint32_t pair_bytes() { return WIDE ? blockDim.x * 16u : (uint32_t)BLOCK_THREADS * 16u; }
    __device__ __forceinline__ static uint32_t mine(uint32_t smem32) { return smem32 + threadIdx.x * (P == 1 ? 8u : 16u); }
    __device__ __forceinline__ static void load(uint32_t my, uint32_t off, Pt<P> &o) {
        const uint32_t p = my + off;
        if constexpr (P == 1) {
            lds64(p, o.v[0]);
        } else {
#pragma unroll
            for (int g = 0; g < P / 2; ++g) lds128(p + g * pair_bytes(), o.v[2 * g], o.v[2 * g + 1]);
        }
    }
    __device__ __forceinline__ static void store(uint32_t my, uint32_t off, const Pt<P> &o) {
        const uint32_t p = my + off;
        if constexpr (P == 1) {
            sts64(p, o.v[0]);
        } else {
#pragma unroll
            for (int g = 0; g < P / 2; ++g) sts128(p + g * pair_bytes(), o.v[2 * g], o.v[2 * g + 1]);
        }
    }
};

// destination store of a micro-instruction: predicated, not branched (d < 0 = result stays in ACC only)
template <int P, bool WIDE = false>
__device__ __forceinline__ void store_if(uint32_t my, int32_t fd, const Pt<P> &o) {
    const uint32_t p = my + (uint32_t)fd;
#ifdef SAF_CHECKED
    if (fd >= 0) {
        if constexpr (P == 1) { SAF_CHECK("store_if", p, 8u); }
        else {
            using RF_ = Rf<P, WIDE>;
            for (int g = 0; g < P / 2; ++g) SAF_CHECK("store_if", p + g * RF_::pair_bytes(), 16u);
        }
    }
#endif
    if constexpr (WIDE && P == 4) {
        const uint32_t p2 = p + Rf<P, WIDE>::pair_bytes();
        asm volatile("{ .reg .pred q; setp.ge.s32 q, %0, 0; @q st.shared.v2.f64 [%1], {%3, %4}; @q st.shared.v2.f64 [%2], {%5, %6}; }" ::"r"(fd),
                     "r"(p), "r"(p2), "d"(o.v[0]), "d"(o.v[1]), "d"(o.v[2]), "d"(o.v[3])
                     : "memory");
    } else if constexpr (P == 1) {
        asm volatile("{ .reg .pred q; setp.ge.s32 q, %0, 0; @q st.shared.f64 [%1], %2; }" ::"r"(fd), "r"(p), "d"(o.v[0]) : "memory");
    } else if constexpr (P == 2) {
        asm volatile("{ .reg .pred q; setp.ge.s32 q, %0, 0; @q st.shared.v2.f64 [%1], {%2, %3}; }" ::"r"(fd), "r"(p), "d"(o.v[0]),
                     "d"(o.v[1])
                     : "memory");
    } else if constexpr (P == 4) {
        asm volatile("{ .reg .pred q; setp.ge.s32 q, %0, 0; @q st.shared.v2.f64 [%1], {%2, %3}; @q st.shared.v2.f64 [%1+2048], {%4, %5}; }" ::"r"(fd),
                     "r"(p), "d"(o.v[0]), "d"(o.v[1]), "d"(o.v[2]), "d"(o.v[3])
                     : "memory");
    } else {
        static_assert(P == 8, "points per thread");
        asm volatile("{ .reg .pred q; setp.ge.s32 q, %0, 0; @q st.shared.v2.f64 [%1], {%2, %3}; @q st.shared.v2.f64 [%1+2048], {%4, %5}; "
                     "@q st.shared.v2.f64 [%1+4096], {%6, %7}; @q st.shared.v2.f64 [%1+6144], {%8, %9}; }" ::"r"(fd),
                     "r"(p), "d"(o.v[0]), "d"(o.v[1]), "d"(o.v[2]), "d"(o.v[3]), "d"(o.v[4]), "d"(o.v[5]), "d"(o.v[6]), "d"(o.v[7])
                     : "memory");
    }
}
static_assert(BLOCK_THREADS * 16 == 2048, "store_if hard-codes the pair stride of 128-thread blocks");

#define SAF_FOR_P _Pragma("unroll") for (int p = 0; p < P; ++p)

// Markers for tools/patch_brx.py: nvcc lowers a C++ switch to a compare tree (ptxas then folds only 3-case
// leaves into jump tables); the build rewrites the PTX of hot_run to ONE brx.idx over all micro-ops.  The
// markers are PTX comments that carry the opcode register / the case value to the patcher.
#define SAF_DISPATCH(u) asm volatile("// SAF_DISPATCH %0 %1;" ::"r"(u), "n"((int)U_COUNT));
#define SAF_DISPATCH_N(u, n) asm volatile("// SAF_DISPATCH %0 %1;" ::"r"(u), "n"((int)(n)));
#define SAF_CASE(v) asm volatile("// SAF_CASE %0;" ::"n"((int)(v)));
#define SAF_CASE_DEFAULT asm volatile("// SAF_CASE default;");
// end of a handler: the next dispatch happens right here (threaded code) -- the patcher replaces the jump back to
// the loop head that follows this marker by its own brx.idx; `id` keeps the compiler from merging handler tails
#define SAF_REDISPATCH(u, id) asm volatile("// SAF_REDISPATCH %0 %1;" ::"r"(u), "n"((int)(id)));

__device__ __forceinline__ uint2 split64(unsigned long long w) { return make_uint2((uint32_t)w, (uint32_t)(w >> 32)); }

} // namespace

// Points of a thread.  P == 1: i = tile_base + tid.  P >= 2: pair g covers points
// tile_base + g * 2 * BLOCK + 2 * tid + {0, 1}: every 128-bit access of a warp is fully coalesced.
template <int P, bool WIDE = false>
__device__ __forceinline__ unsigned long long point_index(unsigned long long tile_base, int p) {
    if constexpr (P == 1) return tile_base + threadIdx.x;
    else return tile_base + (unsigned long long)(p >> 1) * (2 * (WIDE ? blockDim.x : (unsigned)BLOCK_THREADS)) + 2 * threadIdx.x + (p & 1);
}

// ---- cold micro-ops ---------------------------------------------------------------------------------
// pow, powi, 1/expm1, the builtins (61 FnOps of devmath.cuh), operand staging and the rare operand-kind
// combinations: decoded at run time, out of line, memory to memory.  ACC is read from / written to the
// reserved slot `acc_off`; hot_run stores ACC there before it returns and reloads it when re-entered.
// The special functions themselves are ONE out-of-line copy each for the whole module (every layout's cold_op and
// every entry point calls the same code): inlined per layout and per point they made up most of a 20 MB image.
__device__ __noinline__ double cold_builtin1(uint32_t fn, double a) { return dm::builtin1(fn, a).v; }
__device__ __noinline__ double cold_builtin2(uint32_t fn, double a, double b) { return dm::builtin2(fn, a, b).v; }
__device__ __noinline__ double cold_builtin3(uint32_t fn, double a, double b, double c) { return dm::builtin3(fn, a, b, c).v; }
__device__ __noinline__ double cold_builtin4(uint32_t fn, double a, double b, double c, double d) {
    return dm::builtin4(fn, a, b, c, d).v;
}
__device__ __noinline__ double cold_pow(double a, double b) { return pow(a, b); }
__device__ __noinline__ double cold_powi(double a, int n) { return dm::powi(a, n).v; }
__device__ __noinline__ double cold_recip_expm1(double a) { return __ddiv_rn(1.0, expm1(a)); }

template <int P, bool RM = false, bool WIDE = false>
__device__ __noinline__ void cold_op(uint32_t rf32, const unsigned char *img, uint32_t cur, int32_t acc_off, int32_t x0_off,
                                     int32_t x1_off) {
    using RF = Rf<P, WIDE && P == 4>;
    const uint32_t my = RF::mine(rf32);
    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(img + cur);
    const uint2 q0 = split64(w[0]), q1 = split64(w[1]), q2 = split64(w[2]), q3 = split64(w[3]);
    // word 0 holds the NEXT opcode; register-machine images (ruop.h) carry the cold micro-op above the kinds
    const uint32_t uop = RM ? (q3.x >> 8) : q3.y, fa = q1.x, fb = q1.y, fc = q2.x, imm = q2.y, kinds = q3.x & 0xffu;
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
        SAF_FOR_P r.v[p] = cold_pow(a.v[p], b.v[p]);
        break;
    case U_G + G_DIV:
        any(kinds & 3u, fa, a); any((kinds >> 2) & 3u, fb, b);
        SAF_FOR_P r.v[p] = __ddiv_rn(a.v[p], b.v[p]);
        break;
    case U_G + G_POWI:
        any(kinds & 3u, fa, a);
        SAF_FOR_P r.v[p] = cold_powi(a.v[p], (int)imm);
        break;
    case U_G + G_RECIPEXPM1:
        any(kinds & 3u, fa, a);
        if (fb != NO_PREFIX) {
            const double k1 = konst(fb), k2 = konst(fb + 8);
            SAF_FOR_P a.v[p] = __fma_rn(a.v[p], k1, k2);
        }
        SAF_FOR_P r.v[p] = cold_recip_expm1(a.v[p]);
        break;
    case U_G + G_BUILTIN1:
        any(kinds & 3u, fa, a);
        SAF_FOR_P r.v[p] = cold_builtin1(imm, a.v[p]);
        break;
    case U_G + G_BUILTIN2:
        any(kinds & 3u, fa, a); any((kinds >> 2) & 3u, fb, b);
        SAF_FOR_P r.v[p] = cold_builtin2(imm, a.v[p], b.v[p]);
        break;
    case U_G + G_BUILTIN3: {
        any(kinds & 3u, fa, a);
        Pt<P> x0;
        RF::load(my, (uint32_t)x0_off, x0);
        RF::load(my, (uint32_t)x1_off, c);
        SAF_FOR_P r.v[p] = cold_builtin3(imm, x0.v[p], c.v[p], a.v[p]);
    } break;
    case U_G + G_BUILTIN4: {
        any(kinds & 3u, fa, a); any((kinds >> 2) & 3u, fb, b);
        Pt<P> x0;
        RF::load(my, (uint32_t)x0_off, x0);
        RF::load(my, (uint32_t)x1_off, c);
        SAF_FOR_P r.v[p] = cold_builtin4(imm, x0.v[p], c.v[p], a.v[p], b.v[p]);
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
// Runs micro-instructions from code offset `pc` -- over as many iterations of the program as remain -- until the
// program ends for the last time (returns HOT_DONE) or a cold micro-op is met (returns its code offset in the low
// word, the iteration counter in the high word; ACC has been written to its reserved slot).  The code is executed
// out of the shared-memory window at `win_s` (CODE_WINDOW_BYTES, uop.h); the window that contains `pc` must be
// staged when hot_run is entered.  K_SMEM: the constant table sits in shared memory at `k_s`, otherwise it is
// read from the image in global memory.  All threads of the block run hot_run together (same program, same
// iteration count), so the window restaging may use block barriers.
constexpr unsigned long long HOT_DONE = ~0ull;

__device__ __forceinline__ void stage_window(uint32_t win_sm, const unsigned char *code_g, uint32_t win_g, uint32_t code_bytes) {
    const uint32_t bytes = min(CODE_WINDOW_BYTES, code_bytes - win_g);
    const uint4 *src = reinterpret_cast<const uint4 *>(code_g + win_g);
    __syncthreads(); // everybody is done with the previous window
    for (uint32_t i = threadIdx.x; i < bytes / 16u; i += blockDim.x) {
        const uint4 v = __ldg(src + i);
        SAF_CHECK("stage_window", win_sm + 16u * i, 16u);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(win_sm + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    __syncthreads();
}

template <int P, bool K_SMEM, bool WIDE4 = false>
__device__ __noinline__ unsigned long long hot_run(uint32_t smem32, uint32_t ctx_off, const unsigned char *img_g,
                                                   uint32_t pc_code, uint32_t it, uint32_t iters) {
    static_assert(!WIDE4 || P == 4, "the wide layouts of one / two points per thread need no pair stride");
    using RF = Rf<P, WIDE4>;
    const uint32_t my = RF::mine(smem32);
    const uint32_t ctx = smem32 + ctx_off;                           // BlockCtx
    const uint32_t k_sm = ctx + (uint32_t)sizeof(BlockCtx);           // constant table (K_SMEM)
    const uint32_t acc_off = lds_u32(ctx + (uint32_t)offsetof(BlockCtx, acc_off));
    const uint32_t code_off = lds_u32(ctx + (uint32_t)offsetof(BlockCtx, code_off));
    const uint32_t code_bytes = lds_u32(ctx + (uint32_t)offsetof(BlockCtx, code_bytes));
    const uint32_t win_sm = k_sm + (K_SMEM ? code_off : 0u);          // code window
    const unsigned char *const code_g = img_g + code_off;
    uint32_t win_g = pc_code & ~(CODE_WINDOW_BYTES - 1u);             // code offset of the staged window
    uint32_t pc = win_sm + (pc_code & (CODE_WINDOW_BYTES - 1u));      // shared address of the next instruction

    auto img32 = [&](uint32_t addr) -> uint32_t { return lds_u32(addr); }; // word of the current instruction
    auto konst = [&](uint32_t off) -> double {
        if constexpr (K_SMEM) {
            double k;
            lds64(k_sm + off, k);
            return k;
        } else {
            return __ldg(reinterpret_cast<const double *>(img_g + off));
        }
    };

    Pt<P> acc;
    RF::load(my, acc_off, acc);

    // Every handler ends with "store the destination, fetch the next micro-instruction" (SAF_NEXT).  With
    // THREADED dispatch it then jumps through the table itself: one indirect branch per micro-op and no trip
    // through a common loop head -- shorter latency per micro-op, which pays when few warps are resident (small
    // batches, big register files: the P = 1 and P = 2 layouts; double pendulum -7 %, terrain gradient -13 %).
    // Otherwise all handlers return to one dispatch at the loop head, which ptxas keeps entirely in the uniform
    // datapath: fewer issue slots, which pays when the machine is full (P >= 4; Clifford +5 %, Aizawa +15 %
    // against threaded).  The opcode travels one instruction ahead of its operands (uop.h): it is read from the
    // instruction itself only when the loop is entered.
#ifndef SAF_THREADED_MAXP
#define SAF_THREADED_MAXP 2
#endif
    constexpr bool THREADED = P <= SAF_THREADED_MAXP;
    uint32_t uop = lds_u32(pc + 28);
    int32_t fd;
    uint32_t fa, fb, cur, uop_now;
#define SAF_FETCH                                                                                          \
    {                                                                                                      \
        const uint4 q_ = lds_u128(pc); /* {next uop, d, a, b} */                                           \
        fd = (int32_t)q_.y; fa = q_.z; fb = q_.w;                                                          \
        cur = pc; pc += UINSTR_BYTES;                                                                      \
        uop_now = uop; uop = q_.x;                                                                         \
    }
#define SAF_ROTATE /* the prefetched instruction becomes the current one */                               \
    {                                                                                                      \
        fd = (int32_t)qn.y; fa = qn.z; fb = qn.w;                                                          \
        cur = pc; pc += UINSTR_BYTES;                                                                      \
        uop_now = uop; uop = qn.x;                                                                         \
    }
#define SAF_NEXT(id)                                                                                       \
    {                                                                                                      \
        if constexpr (THREADED) {                                                                          \
            if (fd >= 0) RF::store(my, (uint32_t)fd, acc);                                                 \
            SAF_ROTATE SAF_REDISPATCH(uop_now, id) continue;                                               \
        } else {                                                                                           \
            break; /* common tail below the switch: store, then the fetch at the loop head */              \
        }                                                                                                  \
    }
#define SAF_RESTART(id) /* after U_END / U_NEXTWIN (pc has changed): no destination */                     \
    {                                                                                                      \
        if constexpr (THREADED) { qn = lds_u128(pc); SAF_ROTATE SAF_REDISPATCH(uop_now, id) }              \
        continue;                                                                                          \
    }
    // threaded handlers fetch the NEXT micro-instruction as their first action (pc already points at it), so that
    // its latency overlaps their own operand loads
#define SAF_ENTER(v) SAF_CASE(v) if constexpr (THREADED) { qn = lds_u128(pc); }
    uint4 qn = make_uint4(0u, 0u, 0u, 0u);
    if constexpr (THREADED) { qn = lds_u128(pc); SAF_ROTATE }
    for (;;) {
            if constexpr (!THREADED) SAF_FETCH
            Pt<P> a, b, c;

#define LS(x, f) RF::load(my, (f), x);
#define LK(x, f) { const double k_ = konst(f); SAF_FOR_P x.v[p] = k_; }
#define LA(x) x = acc;
#define FC (img32(cur + 16)) /* third operand offset */

// light unary: a from a slot / from ACC
#define U1CASE(OP, EXPR)                                                                                   \
    case U_U1 + 2 * ((OP) - D_FIRST_LIGHT_UNARY): SAF_ENTER(U_U1 + 2 * ((OP) - D_FIRST_LIGHT_UNARY)) LS(a, fa) SAF_FOR_P { const double x = a.v[p]; acc.v[p] = (EXPR); } SAF_NEXT(U_U1 + 2 * ((OP) - D_FIRST_LIGHT_UNARY)) \
    case U_U1 + 2 * ((OP) - D_FIRST_LIGHT_UNARY) + 1: SAF_ENTER(U_U1 + 2 * ((OP) - D_FIRST_LIGHT_UNARY) + 1) SAF_FOR_P { const double x = acc.v[p]; acc.v[p] = (EXPR); } SAF_NEXT(U_U1 + 2 * ((OP) - D_FIRST_LIGHT_UNARY) + 1)
// heavy unary: operand from a slot, {plain, affine prefix}; BODY maps a.v -> acc.v and never reads ACC
#define H1PREFIX { const double k1_ = konst(fb), k2_ = konst(fb + 8); SAF_FOR_P a.v[p] = __fma_rn(a.v[p], k1_, k2_); }
#define H1CASE(H, BODY)                                                                                    \
    case U_H1 + 2 * (H) + 0: SAF_ENTER(U_H1 + 2 * (H) + 0) LS(a, fa) { BODY } SAF_NEXT(U_H1 + 2 * (H) + 0)                                                     \
    case U_H1 + 2 * (H) + 1: SAF_ENTER(U_H1 + 2 * (H) + 1) LS(a, fa) H1PREFIX { BODY } SAF_NEXT(U_H1 + 2 * (H) + 1)
// commutative binary: SS SK SA KA AA
#define BCCASE(I, EXPR)                                                                                    \
    case U_BC + 5 * (I) + BC_SS: SAF_ENTER(U_BC + 5 * (I) + BC_SS) LS(a, fa) LS(b, fb) SAF_FOR_P { const double x = a.v[p], y = b.v[p]; acc.v[p] = (EXPR); } SAF_NEXT(U_BC + 5 * (I) + BC_SS) \
    case U_BC + 5 * (I) + BC_SK: SAF_ENTER(U_BC + 5 * (I) + BC_SK) LS(a, fa) { const double y = konst(fb); SAF_FOR_P { const double x = a.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_BC + 5 * (I) + BC_SK) \
    case U_BC + 5 * (I) + BC_SA: SAF_ENTER(U_BC + 5 * (I) + BC_SA) LS(a, fa) SAF_FOR_P { const double x = a.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } SAF_NEXT(U_BC + 5 * (I) + BC_SA) \
    case U_BC + 5 * (I) + BC_KA: SAF_ENTER(U_BC + 5 * (I) + BC_KA) { const double x = konst(fa); SAF_FOR_P { const double y = acc.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_BC + 5 * (I) + BC_KA) \
    case U_BC + 5 * (I) + BC_AA: SAF_ENTER(U_BC + 5 * (I) + BC_AA) SAF_FOR_P { const double x = acc.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } SAF_NEXT(U_BC + 5 * (I) + BC_AA)
// ordered binary: SS SK KS SA AS KA AK AA
#define BNCASE(I, EXPR)                                                                                    \
    case U_BN + 8 * (I) + BN_SS: SAF_ENTER(U_BN + 8 * (I) + BN_SS) LS(a, fa) LS(b, fb) SAF_FOR_P { const double x = a.v[p], y = b.v[p]; acc.v[p] = (EXPR); } SAF_NEXT(U_BN + 8 * (I) + BN_SS) \
    case U_BN + 8 * (I) + BN_SK: SAF_ENTER(U_BN + 8 * (I) + BN_SK) LS(a, fa) { const double y = konst(fb); SAF_FOR_P { const double x = a.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_BN + 8 * (I) + BN_SK) \
    case U_BN + 8 * (I) + BN_KS: SAF_ENTER(U_BN + 8 * (I) + BN_KS) LS(b, fb) { const double x = konst(fa); SAF_FOR_P { const double y = b.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_BN + 8 * (I) + BN_KS) \
    case U_BN + 8 * (I) + BN_SA: SAF_ENTER(U_BN + 8 * (I) + BN_SA) LS(a, fa) SAF_FOR_P { const double x = a.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } SAF_NEXT(U_BN + 8 * (I) + BN_SA) \
    case U_BN + 8 * (I) + BN_AS: SAF_ENTER(U_BN + 8 * (I) + BN_AS) LS(b, fb) SAF_FOR_P { const double x = acc.v[p], y = b.v[p]; acc.v[p] = (EXPR); } SAF_NEXT(U_BN + 8 * (I) + BN_AS) \
    case U_BN + 8 * (I) + BN_KA: SAF_ENTER(U_BN + 8 * (I) + BN_KA) { const double x = konst(fa); SAF_FOR_P { const double y = acc.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_BN + 8 * (I) + BN_KA) \
    case U_BN + 8 * (I) + BN_AK: SAF_ENTER(U_BN + 8 * (I) + BN_AK) { const double y = konst(fb); SAF_FOR_P { const double x = acc.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_BN + 8 * (I) + BN_AK) \
    case U_BN + 8 * (I) + BN_AA: SAF_ENTER(U_BN + 8 * (I) + BN_AA) SAF_FOR_P { const double x = acc.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } SAF_NEXT(U_BN + 8 * (I) + BN_AA)
// fused multiply-add family: x * y (+/-) z
#define TCASE(T, EXPR)                                                                                     \
    case U_T + 10 * (T) + T_SSS: SAF_ENTER(U_T + 10 * (T) + T_SSS) { const uint32_t fc = FC; LS(a, fa) LS(b, fb) LS(c, fc) SAF_FOR_P { const double x = a.v[p], y = b.v[p], z = c.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_SSS) \
    case U_T + 10 * (T) + T_SSK: SAF_ENTER(U_T + 10 * (T) + T_SSK) { const uint32_t fc = FC; LS(a, fa) LS(b, fb) const double z = konst(fc); SAF_FOR_P { const double x = a.v[p], y = b.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_SSK) \
    case U_T + 10 * (T) + T_SSA: SAF_ENTER(U_T + 10 * (T) + T_SSA) { LS(a, fa) LS(b, fb) SAF_FOR_P { const double x = a.v[p], y = b.v[p], z = acc.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_SSA) \
    case U_T + 10 * (T) + T_SKS: SAF_ENTER(U_T + 10 * (T) + T_SKS) { const uint32_t fc = FC; LS(a, fa) LS(c, fc) const double y = konst(fb); SAF_FOR_P { const double x = a.v[p], z = c.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_SKS) \
    case U_T + 10 * (T) + T_SKK: SAF_ENTER(U_T + 10 * (T) + T_SKK) { const uint32_t fc = FC; LS(a, fa) const double y = konst(fb), z = konst(fc); SAF_FOR_P { const double x = a.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_SKK) \
    case U_T + 10 * (T) + T_SKA: SAF_ENTER(U_T + 10 * (T) + T_SKA) { LS(a, fa) const double y = konst(fb); SAF_FOR_P { const double x = a.v[p], z = acc.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_SKA) \
    case U_T + 10 * (T) + T_SAS: SAF_ENTER(U_T + 10 * (T) + T_SAS) { const uint32_t fc = FC; LS(a, fa) LS(c, fc) SAF_FOR_P { const double x = a.v[p], y = acc.v[p], z = c.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_SAS) \
    case U_T + 10 * (T) + T_SAK: SAF_ENTER(U_T + 10 * (T) + T_SAK) { const uint32_t fc = FC; LS(a, fa) const double z = konst(fc); SAF_FOR_P { const double x = a.v[p], y = acc.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_SAK) \
    case U_T + 10 * (T) + T_KAS: SAF_ENTER(U_T + 10 * (T) + T_KAS) { const uint32_t fc = FC; LS(c, fc) const double x = konst(fa); SAF_FOR_P { const double y = acc.v[p], z = c.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_KAS) \
    case U_T + 10 * (T) + T_KAK: SAF_ENTER(U_T + 10 * (T) + T_KAK) { const uint32_t fc = FC; const double x = konst(fa), z = konst(fc); SAF_FOR_P { const double y = acc.v[p]; acc.v[p] = (EXPR); } } SAF_NEXT(U_T + 10 * (T) + T_KAK)

// fused chains (uop.h U_FM / U_FQ): LOADS fetches the operands of the multiply, PROD is the product expression
#define FI (img32(cur + 20)) /* imm field: slot operand of an ADDS suffix */
#define FMCASE(V, LOADS, PROD)                                                                             \
    case U_FM + 3 * (V) + F_MULK: SAF_ENTER(U_FM + 3 * (V) + F_MULK) { const double k_ = konst(FC); LOADS SAF_FOR_P { acc.v[p] = __dmul_rn(k_, (PROD)); } } SAF_NEXT(U_FM + 3 * (V) + F_MULK) \
    case U_FM + 3 * (V) + F_ADDS: SAF_ENTER(U_FM + 3 * (V) + F_ADDS) { const uint32_t fs = FI; LS(c, fs) LOADS SAF_FOR_P { acc.v[p] = __dadd_rn(c.v[p], (PROD)); } } SAF_NEXT(U_FM + 3 * (V) + F_ADDS) \
    case U_FM + 3 * (V) + F_MULK_ADDS: SAF_ENTER(U_FM + 3 * (V) + F_MULK_ADDS) { const uint32_t fs = FI; const double k_ = konst(FC); LS(c, fs) LOADS SAF_FOR_P { acc.v[p] = __dadd_rn(c.v[p], __dmul_rn(k_, (PROD))); } } SAF_NEXT(U_FM + 3 * (V) + F_MULK_ADDS)
#define FQCASE(Q, LOADS, SRC)                                                                              \
    case U_FQ + 2 * (Q) + F_MULK: SAF_ENTER(U_FQ + 2 * (Q) + F_MULK) { const double k_ = konst(FC); LOADS SAF_FOR_P { const double x = SRC.v[p]; acc.v[p] = __dmul_rn(k_, __dmul_rn(x, x)); } } SAF_NEXT(U_FQ + 2 * (Q) + F_MULK) \
    case U_FQ + 2 * (Q) + F_ADDS: SAF_ENTER(U_FQ + 2 * (Q) + F_ADDS) { const uint32_t fs = FI; LS(c, fs) LOADS SAF_FOR_P { const double x = SRC.v[p]; acc.v[p] = __dadd_rn(c.v[p], __dmul_rn(x, x)); } } SAF_NEXT(U_FQ + 2 * (Q) + F_ADDS)

// t = S[a] + K[b], kept in the slot named by the kinds word when that is >= 0, then squared / scaled twice (uop.h U_FS)
#define FSHEAD const int32_t d1_ = (int32_t)img32(cur + 24); LS(a, fa) { const double y_ = konst(fb); SAF_FOR_P a.v[p] = __dadd_rn(a.v[p], y_); } store_if<P, WIDE4>(my, d1_, a);

// sin / cos + "S[imm] + ACC" / "fma(S[imm], K[c], ACC)" (uop.h U_FT; P == 4 only, other layouts never receive them)
#define FTBODY(TRIG, PRE, TAIL) if constexpr (P == 4) { const uint32_t fs = FI; LS(a, fa) LS(c, fs) PRE Pt<P> t_; TRIG(a.v, t_.v); TAIL }
#define FTCASE(T, TRIG)                                                                                    \
    case U_FT + 4 * (T) + 0: SAF_ENTER(U_FT + 4 * (T) + 0) { FTBODY(TRIG, , SAF_FOR_P acc.v[p] = __dadd_rn(c.v[p], t_.v[p]);) } SAF_NEXT(U_FT + 4 * (T) + 0) \
    case U_FT + 4 * (T) + 1: SAF_ENTER(U_FT + 4 * (T) + 1) { FTBODY(TRIG, H1PREFIX, SAF_FOR_P acc.v[p] = __dadd_rn(c.v[p], t_.v[p]);) } SAF_NEXT(U_FT + 4 * (T) + 1) \
    case U_FT + 4 * (T) + 2: SAF_ENTER(U_FT + 4 * (T) + 2) { FTBODY(TRIG, , const double k_ = konst(FC); SAF_FOR_P acc.v[p] = __fma_rn(c.v[p], k_, t_.v[p]);) } SAF_NEXT(U_FT + 4 * (T) + 2) \
    case U_FT + 4 * (T) + 3: SAF_ENTER(U_FT + 4 * (T) + 3) { FTBODY(TRIG, H1PREFIX, const double k_ = konst(FC); SAF_FOR_P acc.v[p] = __fma_rn(c.v[p], k_, t_.v[p]);) } SAF_NEXT(U_FT + 4 * (T) + 3)

// trig1(S[a] prefix b) (+ | fma K[c]) trig2(S[imm] prefix kinds-word)  (uop.h U_FD; P == 4 only)
#define FDBODY(TRIG1, TRIG2, TAIL) if constexpr (P == 4) {                                                 \
        const uint32_t f2_ = FI, pb2_ = img32(cur + 24);                                                   \
        LS(a, fa) LS(b, f2_)                                                                               \
        { const double k1_ = konst(fb), k2_ = konst(fb + 8); SAF_FOR_P a.v[p] = __fma_rn(a.v[p], k1_, k2_); } \
        { const double k1_ = konst(pb2_), k2_ = konst(pb2_ + 8); SAF_FOR_P b.v[p] = __fma_rn(b.v[p], k1_, k2_); } \
        Pt<P> u_, v_; TRIG1(a.v, u_.v); TRIG2(b.v, v_.v); TAIL }
#define FDCASE(T1, T2, TRIG1, TRIG2)                                                                       \
    case U_FD + 4 * (T1) + 2 * (T2) + 0: SAF_ENTER(U_FD + 4 * (T1) + 2 * (T2) + 0) { FDBODY(TRIG1, TRIG2, SAF_FOR_P acc.v[p] = __dadd_rn(u_.v[p], v_.v[p]);) } SAF_NEXT(U_FD + 4 * (T1) + 2 * (T2) + 0) \
    case U_FD + 4 * (T1) + 2 * (T2) + 1: SAF_ENTER(U_FD + 4 * (T1) + 2 * (T2) + 1) { FDBODY(TRIG1, TRIG2, const double k_ = konst(FC); SAF_FOR_P acc.v[p] = __fma_rn(u_.v[p], k_, v_.v[p]);) } SAF_NEXT(U_FD + 4 * (T1) + 2 * (T2) + 1)

            SAF_DISPATCH(uop_now)
            switch (uop_now) {
            case U_END: SAF_CASE(U_END) { // a = code offset where the next iteration starts
                if (++it >= iters) return HOT_DONE;
                const uint32_t w = fa & ~(CODE_WINDOW_BYTES - 1u);
                if (w != win_g) {
                    win_g = w;
                    stage_window(win_sm, code_g, win_g, code_bytes);
                }
                pc = win_sm + (fa & (CODE_WINDOW_BYTES - 1u));
                // in-kernel feedback for images that are not double buffered: parameter j <- result j (result
                // slots never alias parameter slots, lower.cpp); the state stays on chip between iterations
                const uint32_t n_fb = lds_u32(ctx + (uint32_t)offsetof(BlockCtx, n_params));
                if (n_fb != 0u) {
                    const uint32_t k_mask = lds_u32(ctx + (uint32_t)offsetof(BlockCtx, result_k_mask));
                    for (uint32_t j = 0; j < n_fb; ++j) {
                        const int32_t s_ = (int32_t)lds_u32(ctx + (uint32_t)offsetof(BlockCtx, param_off) + 4u * j);
                        if (s_ < 0) continue;
                        const uint32_t r_ = lds_u32(ctx + (uint32_t)offsetof(BlockCtx, result_off) + 4u * j);
                        Pt<P> v;
                        if ((k_mask >> j) & 1u) {
                            const double k = konst(r_);
                            SAF_FOR_P v.v[p] = k;
                        } else {
                            RF::load(my, r_, v);
                        }
                        RF::store(my, (uint32_t)s_, v);
                    }
                }
                SAF_RESTART(U_END)
            }
            case U_NEXTWIN: SAF_CASE(U_NEXTWIN) // end of the staged window: stage the next one, continue at its start
                win_g += CODE_WINDOW_BYTES;
                stage_window(win_sm, code_g, win_g, code_bytes);
                pc = win_sm;
                SAF_RESTART(U_NEXTWIN)
            case U_LOADK: SAF_ENTER(U_LOADK) { const double k_ = konst(fa); SAF_FOR_P acc.v[p] = k_; } SAF_NEXT(U_LOADK)

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

            FMCASE(BC_SS, LS(a, fa) LS(b, fb), __dmul_rn(a.v[p], b.v[p]))
            FMCASE(BC_SK, LS(a, fa) const double y_ = konst(fb);, __dmul_rn(a.v[p], y_))
            FMCASE(BC_SA, LS(a, fa), __dmul_rn(a.v[p], acc.v[p]))
            FMCASE(BC_KA, const double x_ = konst(fa);, __dmul_rn(x_, acc.v[p]))
            FMCASE(BC_AA, , __dmul_rn(acc.v[p], acc.v[p]))
            FQCASE(0, LS(a, fa), a)
            FQCASE(1, , acc)
            FTCASE(0, lm::sin_n<P>)
            FTCASE(1, lm::cos_n<P>)
            FDCASE(0, 0, lm::sin_n<P>, lm::sin_n<P>)
            FDCASE(0, 1, lm::sin_n<P>, lm::cos_n<P>)
            FDCASE(1, 0, lm::cos_n<P>, lm::sin_n<P>)
            FDCASE(1, 1, lm::cos_n<P>, lm::cos_n<P>)
            case U_FS + FS_SQ: SAF_ENTER(U_FS + FS_SQ) { FSHEAD SAF_FOR_P acc.v[p] = __dmul_rn(a.v[p], a.v[p]); } SAF_NEXT(U_FS + FS_SQ)
            case U_FS + FS_SQ_ADDS: SAF_ENTER(U_FS + FS_SQ_ADDS) { const uint32_t fs = FI; FSHEAD LS(c, fs) /* after the store: S[imm] may be the kept slot */ SAF_FOR_P acc.v[p] = __dadd_rn(c.v[p], __dmul_rn(a.v[p], a.v[p])); } SAF_NEXT(U_FS + FS_SQ_ADDS)
            case U_FS + FS_MULK2: SAF_ENTER(U_FS + FS_MULK2) { const double k1_ = konst(FC), k2_ = konst(FI); FSHEAD SAF_FOR_P acc.v[p] = __dmul_rn(k2_, __dmul_rn(k1_, a.v[p])); } SAF_NEXT(U_FS + FS_MULK2)

            case U_F5: SAF_ENTER(U_F5) {
                const double k1_ = konst(fb), k2_ = konst(FC), k3_ = konst(img32(cur + 24));
                const uint32_t s2_ = FI;
                LS(a, fa) LS(b, s2_) LS(c, (uint32_t)fd)
                SAF_FOR_P acc.v[p] = __dadd_rn(c.v[p], __dmul_rn(k3_, __dmul_rn(b.v[p], __dmul_rn(k2_, __dmul_rn(a.v[p], k1_)))));
            } SAF_NEXT(U_F5)

            TCASE(T_FMA, __fma_rn(x, y, z))
            TCASE(T_FMS, __fma_rn(x, y, -z))
            TCASE(T_FNMA, __fma_rn(-x, y, z))
            TCASE(T_FNMS, __fma_rn(-x, y, -z))

            // cold micro-ops leave the loop: ACC goes to its slot, the caller runs cold_op and re-enters
            default: SAF_CASE_DEFAULT
                RF::store(my, acc_off, acc);
                return ((unsigned long long)it << 32) | (win_g + (cur - win_sm));
            }
            if (fd >= 0) RF::store(my, (uint32_t)fd, acc); // reached by the non-threaded handlers only
    }
}

#ifndef SAF_R_THREADED
#define SAF_R_THREADED false
#endif
// The opt-in register machine (SAF_MACHINE=r) is compiled into a cubin of its own (Makefile: kernel_r.cubin, built
// from this file with -DSAF_BUILD_R) that launch.cu loads only when it is asked for: its 775 handlers made the default
// image 22.7 MB and cold instruction fetch is what a 1 M-point single pass pays for.
#ifdef SAF_BUILD_R
#include "kernel_r.cuh"
#else
template <bool K_SMEM, bool THREADED>
__device__ unsigned long long hot_run_r(uint32_t, uint32_t, const unsigned char *, uint32_t, uint32_t, uint32_t); // never instantiated here
#endif

// ---- the kernel ------------------------------------------------------------------------------------------
// RM: the register machine (ruop.h, kernel_r.cuh) runs the program; its image is packed by pack_r.cpp
// WIDE (P <= 2 only): one block of blockDim.x > BLOCK_THREADS threads per SM -- programs with many live values fit
// two 128-thread blocks per SM at most, and one wide block holds more warps in the same shared memory (the code
// window and the constant table exist once)
// STREAMING (entry points saf_interp_p{2,4}s_ksm, d.n_rf = 2 or 3; a kernel of its own so that the tile loop below does
// not weigh on the register allocation of the classic entry points -- ptxas allocates hot_run's registers against
// its callers' live ranges, and with both loops in one kernel every four-points workload lost 15 %; the host grants it to single passes over full-length aligned columns whose register
// file is small -- the HBM-bound programs): shared memory holds n_rf copies of the register file.  A parameter slot's
// image for one tile is a CONTIGUOUS copy of the column tile (Rf layout, point_index), so one elected thread moves
// tile t + n_rf - 1 into the parameter slots of another copy with 1-D bulk copies (cp.async.bulk, completion on that
// copy's mbarrier) while the block evaluates tile t, and result slots go back to the output columns with bulk stores
// that drain while the next tile runs: column traffic never waits for the interpreter and no register holds column
// data.  A partial last tile takes the classic path.
template <int P, bool K_SMEM, bool RM = false, bool WIDE = false, bool STREAM = false>
__device__ __forceinline__ void interp_body(const LaunchDesc &d) {
    const uint32_t block_threads = WIDE ? blockDim.x : (uint32_t)BLOCK_THREADS;
    using RF = Rf<P, WIDE && P == 4>;
    const uint32_t smem32 = smem_base32();

    // ---- behind the register file(s): block context, constant table (K_SMEM), code window ----
    const uint32_t rf_bytes = d.n_slots * (block_threads * (uint32_t)(P * 8));
    const uint32_t n_rf = STREAM ? d.n_rf : 1u;
    const uint32_t ctx_off = n_rf * rf_bytes;
    const uint32_t k_s = ctx_off + (uint32_t)sizeof(BlockCtx);
    const uint32_t win_s = k_s + (K_SMEM ? d.code_off : 0u);
    {
        BlockCtx *ctx = reinterpret_cast<BlockCtx *>(saf_smem + ctx_off);
        if (threadIdx.x == 0) {
            ctx->n_params = d.n_feedback;
            ctx->result_k_mask = d.result_k_mask;
            ctx->acc_off = d.acc_off;
            ctx->code_off = d.code_off;
            ctx->code_bytes = d.code_bytes;
        }
        if (threadIdx.x < MAX_COLS) ctx->param_off[threadIdx.x] = d.param_off[threadIdx.x];
        if (threadIdx.x < MAX_OUTS) ctx->result_off[threadIdx.x] = d.result_off[threadIdx.x];
        if constexpr (K_SMEM) {
            const uint4 *src = reinterpret_cast<const uint4 *>(d.image_global);
            uint4 *dst = reinterpret_cast<uint4 *>(saf_smem + k_s);
            for (uint32_t i = threadIdx.x; i < d.code_off / 16u; i += block_threads) dst[i] = __ldg(src + i);
        }
        stage_window(smem32 + win_s, d.image_global + d.code_off, 0u, d.code_bytes); // barriers inside
    }
    const bool multi_window = d.code_bytes > CODE_WINDOW_BYTES;
    const uint32_t iters = (uint32_t)d.iters;

    const unsigned long long TILE = (unsigned long long)block_threads * P;
    const unsigned long long n_tiles = (d.n_points + TILE - 1) / TILE;

    // the program on the register file at shared address rf32: hot_run until it ends, cold micro-ops in between
    auto run_program = [&](uint32_t rf32, bool restage) {
        if (restage) stage_window(smem32 + win_s, d.image_global + d.code_off, 0u, d.code_bytes);
        const uint32_t ctx_rel = smem32 + ctx_off - rf32;
        uint32_t pc = 0, it = 0;
        for (;;) {
            unsigned long long st;
            if constexpr (RM) st = hot_run_r<K_SMEM, SAF_R_THREADED>(rf32, ctx_rel, d.image_global, pc, it, iters);
            else st = hot_run<P, K_SMEM, WIDE && P == 4>(rf32, ctx_rel, d.image_global, pc, it, iters);
            if (st == HOT_DONE) break;
            pc = (uint32_t)st;
            it = (uint32_t)(st >> 32);
            cold_op<P, RM, WIDE>(rf32, d.image_global, d.code_off + pc, d.acc_off, d.x0_off, d.x1_off);
            pc += UINSTR_BYTES;
        }
    };

    // one tile the classic way: loads -> parameter slots, program, result slots -> stores
    auto classic_tile = [&](unsigned long long tile, bool restage) {
        const unsigned long long tile_base = tile * TILE;
        const uint32_t my = RF::mine(smem32);
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
                    const unsigned long long i = point_index<P, WIDE>(tile_base, 0);
                    if (i < d.n_points) v.v[0] = ld_stream(col + (i < len ? i : len - 1));
                } else {
#pragma unroll
                    for (int g = 0; g < P / 2; ++g) {
                        const unsigned long long i = point_index<P, WIDE>(tile_base, 2 * g);
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
        run_program(smem32, restage);
        // ---- epilogue: result slots -> output columns ----
        for (uint32_t k = 0; k < d.n_results; ++k) {
            Pt<P> v;
            if ((d.result_k_mask >> k) & 1u) {
                const double kv = __ldg(reinterpret_cast<const double *>(d.image_global + (uint32_t)d.result_off[k]));
                SAF_FOR_P v.v[p] = kv;
            } else {
                RF::load(my, (uint32_t)d.result_off[k], v);
            }
            double *out = d.outs[k];
            if constexpr (P == 1) {
                const unsigned long long i = point_index<P, WIDE>(tile_base, 0);
                if (i < d.n_points) st_stream(out + i, v.v[0]);
            } else {
#pragma unroll
                for (int g = 0; g < P / 2; ++g) {
                    const unsigned long long i = point_index<P, WIDE>(tile_base, 2 * g);
                    if (i + 1 < d.n_points) {
                        st_stream2(out + i, v.v[2 * g], v.v[2 * g + 1]);
                    } else if (i < d.n_points) {
                        st_stream(out + i, v.v[2 * g]);
                    }
                }
            }
        }
    };

    if constexpr (STREAM) {
        static_assert(!RM && !WIDE && P >= 2 && K_SMEM, "streaming entry points: 128-thread blocks, P >= 2, constants in shared memory");
        if (n_rf > 1u) {
            const unsigned long long n_full = d.n_points / TILE; // whole tiles: bulk copies of TILE * 8 bytes
            const uint32_t tile_bytes = (uint32_t)TILE * 8u;
            const uint32_t code_sm = (uint32_t)(d.code_bytes < CODE_WINDOW_BYTES ? d.code_bytes : CODE_WINDOW_BYTES);
            const uint32_t bar0 = smem32 + win_s + ((code_sm + 15u) & ~15u);
            uint32_t n_used = 0;
            for (uint32_t j = 0; j < d.n_params; ++j) n_used += d.param_off[j] >= 0;
            if (threadIdx.x == 0) {
                for (uint32_t b = 0; b < n_rf; ++b) mbar_init(bar0 + 8u * b, 1u);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            auto issue = [&](unsigned long long tile, uint32_t buf) { // elected thread only
                const uint32_t bar = bar0 + 8u * buf;
                if (n_used != 0u) mbar_expect_tx(bar, n_used * tile_bytes);
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
                for (uint32_t j = 0; j < d.n_params; ++j)
                    if (d.param_off[j] >= 0)
                        bulk_load(smem32 + buf * rf_bytes + (uint32_t)d.param_off[j], d.cols[j] + tile * TILE, tile_bytes, bar);
            };
            unsigned long long next = blockIdx.x;
            if (threadIdx.x == 0) { // the tiles of iterations 0 .. n_rf - 1, one per copy
                uint32_t k = 0;
                for (; k < n_rf && next < n_full; ++k, next += gridDim.x) issue(next, k);
            }
            uint32_t it = 0;
            for (unsigned long long tile = blockIdx.x; tile < n_full; tile += gridDim.x, ++it) {
                const uint32_t buf = it % n_rf;
                mbar_wait(bar0 + 8u * buf, (it / n_rf) & 1u);
                run_program(smem32 + buf * rf_bytes, false);
                fence_async_smem(); // result slots written through the generic proxy -> visible to the bulk stores
                // ONE block barrier per tile.  Before it, the elected thread makes sure that the stores of tile
                // it - n_rf + 1 have finished reading their result slots (the program of tile it + 1 overwrites them);
                // behind it every thread is done with this copy's parameter slots and has written its result slots.
                if (threadIdx.x == 0) {
                    if (n_rf == 2u) bulk_wait_read<0>();
                    else bulk_wait_read<1>();
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    for (uint32_t k = 0; k < d.n_results; ++k)
                        bulk_store(d.outs[k] + tile * TILE, smem32 + buf * rf_bytes + (uint32_t)d.result_off[k], tile_bytes);
                    bulk_commit();
                    if (next < n_full) { // the tile of iteration it + n_rf lands in the copy this tile has just left
                        issue(next, buf);
                        next += gridDim.x;
                    }
                }
            }
            if (threadIdx.x == 0) bulk_wait_all(); // shared memory must outlive the stores that read it
            __syncthreads();
            if (n_full < n_tiles && (n_full % gridDim.x) == blockIdx.x) classic_tile(n_full, false);
            return;
        }
    }

    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        classic_tile(tile, multi_window && tile != blockIdx.x); // the previous tile left the last window staged
}

// ---- entry points (extern "C": looked up by name in the cubin, launch.cu) ---------------------------------
#define SAF_ENTRY(NAME, P, SM, MINB)                                                                       \
    extern "C" __global__ void __launch_bounds__(BLOCK_THREADS, MINB) NAME(const __grid_constant__ LaunchDesc d) { \
        interp_body<P, SM>(d);                                                                             \
    }
// *_ksm: constant table staged in shared memory; *_kgm: read from global memory (very large tables)
#ifdef SAF_BUILD_R
#define SAF_ENTRY_R(NAME, SM, MINB)                                                                        \
    extern "C" __global__ void __launch_bounds__(BLOCK_THREADS, MINB) NAME(const __grid_constant__ LaunchDesc d) { \
        interp_body<R_POINTS, SM, true>(d);                                                                \
    }
SAF_ENTRY_R(saf_rinterp_ksm, true, SAF_MINB_R)
#else
SAF_ENTRY(saf_interp_p4_ksm, 4, true, SAF_MINB_P4)
#define SAF_ENTRY_W(NAME, P, SM)                                                                           \
    extern "C" __global__ void __launch_bounds__(WIDE_MAX_THREADS, 1) NAME(const __grid_constant__ LaunchDesc d) { \
        interp_body<P, SM, false, true>(d);                                                                \
    }
#ifndef SAF_FAST_BUILD
SAF_ENTRY(saf_interp_p4_kgm, 4, false, SAF_MINB_P4)
SAF_ENTRY(saf_interp_p2_ksm, 2, true, SAF_MINB_P2)
SAF_ENTRY(saf_interp_p2_kgm, 2, false, SAF_MINB_P2)
SAF_ENTRY_W(saf_interp_p2w_ksm, 2, true)
SAF_ENTRY_W(saf_interp_p2w_kgm, 2, false)
SAF_ENTRY_W(saf_interp_p4w_ksm, 4, true)
SAF_ENTRY_W(saf_interp_p4w_kgm, 4, false)
#define SAF_ENTRY_S(NAME, P, MINB)                                                                         \
    extern "C" __global__ void __launch_bounds__(BLOCK_THREADS, MINB) NAME(const __grid_constant__ LaunchDesc d) { \
        interp_body<P, true, false, false, true>(d);                                                       \
    }
SAF_ENTRY_S(saf_interp_p4s_ksm, 4, SAF_MINB_P4)
SAF_ENTRY_S(saf_interp_p2s_ksm, 2, SAF_MINB_P2)
SAF_ENTRY(saf_interp_p1_ksm, 1, true, SAF_MINB_P1)
SAF_ENTRY(saf_interp_p1_kgm, 1, false, SAF_MINB_P1)
#ifdef SAF_WITH_P8 // eight points per thread: a tuning experiment that lost on every workload; not shipped
SAF_ENTRY(saf_interp_p8_ksm, 8, true, SAF_MINB_P8)
SAF_ENTRY(saf_interp_p8_kgm, 8, false, SAF_MINB_P8)
#endif
#endif // SAF_FAST_BUILD

// ---- FP64 pipe peak probe ---------------------------------------------------------------------------
// The roofline of this path is the FP64 pipe (BASELINE.json north_star), whose peak is not in
// MEASURED_PEAKS.json.  This kernel measures it on the box: 8 independent DFMA chains per thread,
// enough warps to saturate the pipe; bench.py divides lane-instructions by the CUDA-event time.
extern "C" __global__ void __launch_bounds__(256) saf_fp64_peak_kernel(double *sink, int iters, double seed) {
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


// ---- "ideal kernel" probe ---------------------------------------------------------------------------------
// What the lean sin/cos kernels reach WITHOUT an interpreter around them: the Clifford map hard-coded, 4 points
// per thread, state in registers, same launch shape as the interpreter.  bench.py reports it next to the
// interpreter's number: the gap between the two is what interpretation costs, the gap between this and the
// DFMA-chain peak is what the FP64 pipe loses on a realistic instruction mix.
extern "C" __global__ void __launch_bounds__(BLOCK_THREADS, SAF_MINB_P4)
saf_clifford_probe_kernel(double *xs, double *ys, unsigned long long n, int iters, double a, double b, double c, double d) {
    const unsigned long long tile = (unsigned long long)BLOCK_THREADS * 4;
    for (unsigned long long base = (unsigned long long)blockIdx.x * tile; base < n; base += (unsigned long long)gridDim.x * tile) {
        double x[4], y[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const unsigned long long i = base + (unsigned long long)p * BLOCK_THREADS + threadIdx.x;
            x[p] = i < n ? xs[i] : 0.0;
            y[p] = i < n ? ys[i] : 0.0;
        }
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            double t[4], s1[4], c1[4], s2[4], c2[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) t[p] = __dmul_rn(y[p], a);
            lm::sin_n<4>(t, s1);
#pragma unroll
            for (int p = 0; p < 4; ++p) t[p] = __dmul_rn(x[p], a);
            lm::cos_n<4>(t, c1);
#pragma unroll
            for (int p = 0; p < 4; ++p) t[p] = __dmul_rn(x[p], b);
            lm::sin_n<4>(t, s2);
#pragma unroll
            for (int p = 0; p < 4; ++p) t[p] = __dmul_rn(y[p], b);
            lm::cos_n<4>(t, c2);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                x[p] = __fma_rn(c, c1[p], s1[p]);
                y[p] = __fma_rn(d, c2[p], s2[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const unsigned long long i = base + (unsigned long long)p * BLOCK_THREADS + threadIdx.x;
            if (i < n) {
                xs[i] = x[p];
                ys[i] = y[p];
            }
        }
    }
}
#endif // !SAF_BUILD_R

} // namespace saf
