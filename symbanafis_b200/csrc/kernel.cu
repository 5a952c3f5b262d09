// kernel.cu -- the sm_100a interpreter kernel for SymbAnaFis device programs.
//
// One thread evaluates P points (P = 1 or 2).  Every lane of every warp runs the SAME device
// program on different points, so the program counter, the fetched instruction word and the opcode
// switch are warp-uniform: there is no divergence in the dispatch, only inside data-dependent
// special-function loops.
//
//   program words + constants : kernel parameter space = constant bank (ConstBlob), read through the
//                               constant cache with uniform addresses; global-memory fallback
//                               (read-only path) for programs above BLOB_LARGE words
//   live values ("registers")  : shared memory laid out [slot][thread] as f64 (P=1) or f64x2 (P=2):
//                               consecutive lanes touch consecutive 8/16-byte words -> conflict free
//   accumulator                : the result of the previous instruction stays in real registers; the
//                               lowering (lower.cpp) routes single-use values through it so they
//                               never touch shared memory
//   columns                    : struct-of-arrays f64, one coalesced 128-bit (P=2) or 64-bit load per
//                               column per thread, streaming cache hints; results likewise
//
// Arithmetic handlers use the explicit-rounding intrinsics (__dadd_rn, __dmul_rn, __fma_rn, ...) so
// that nvcc can never contract a Mul followed by an Add: the reference distinguishes Mul;Add from
// MulAdd (macros.rs:69-75 vs :128-138).
#include "kernel.cuh"

#include <cstdio>

#include "devmath.cuh"

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

template <int P>
__device__ __forceinline__ void slot_load(const double *base, uint32_t stride, uint32_t idx, Pt<P> &o) {
    const double *p = base + (size_t)idx * stride;
    if constexpr (P == 2) {
        double2 t = *reinterpret_cast<const double2 *>(p);
        o.v[0] = t.x;
        o.v[1] = t.y;
    } else {
        o.v[0] = *p;
    }
}
template <int P>
__device__ __forceinline__ void slot_store(double *base, uint32_t stride, uint32_t idx, const Pt<P> &o) {
    double *p = base + (size_t)idx * stride;
    if constexpr (P == 2) {
        *reinterpret_cast<double2 *>(p) = make_double2(o.v[0], o.v[1]);
    } else {
        *p = o.v[0];
    }
}

#define SAF_FOR_P _Pragma("unroll") for (int p = 0; p < P; ++p)

} // namespace

// WORDS > 0: program in the constant bank; WORDS == 0: program in global memory.
template <int P, int WORDS>
__global__ void __launch_bounds__(256, 2)
saf_interp_kernel(const __grid_constant__ LaunchDesc d,
                  const __grid_constant__ ConstBlob<(WORDS > 0 ? WORDS : 1)> blob) {
    extern __shared__ __align__(16) unsigned char saf_smem[];
    const uint32_t stride = blockDim.x * P;                       // doubles between consecutive slots
    double *const my = reinterpret_cast<double *>(saf_smem) + threadIdx.x * P;

    auto fetch_word = [&](uint32_t pc) -> uint64_t {
        if constexpr (WORDS > 0) return blob.w[pc];
        else return __ldg(reinterpret_cast<const unsigned long long *>(d.code_global) + pc);
    };
    auto fetch_const = [&](uint32_t idx) -> double {
        if constexpr (WORDS > 0) {
            // constants ride in the blob right after the code unless they did not fit
            if (d.consts_global == nullptr) return __longlong_as_double((long long)blob.w[d.n_code + idx]);
        }
        return __ldg(d.consts_global + idx);
    };

    const unsigned long long tile_points = (unsigned long long)blockDim.x * P;
    const unsigned long long n_tiles = (d.n_points + tile_points - 1) / tile_points;

    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long i0 = tile * tile_points + (unsigned long long)threadIdx.x * P;

        // ---- prologue: columns -> parameter slots (scalar.rs:182-207 rules) ----
        for (uint32_t j = 0; j < d.n_params; ++j) {
            const uint32_t s = d.param_slot[j];
            if (s == 0xffffu) continue;
            const double *col = d.cols[j];
            const unsigned long long len = d.col_len[j];
            Pt<P> v;
            SAF_FOR_P v.v[p] = 0.0; // missing column reads as 0.0
            if (col != nullptr && len != 0ull) {
                if (P == 2 && i0 + 1 < len && i0 + 1 < d.n_points) {
                    ld_stream2(col + i0, v.v[0], v.v[P - 1]);
                } else {
                    SAF_FOR_P {
                        const unsigned long long i = i0 + p;
                        if (i < d.n_points) v.v[p] = ld_stream(col + (i < len ? i : len - 1));
                    }
                }
            }
            slot_store<P>(my, stride, s, v);
        }

        Pt<P> acc, x0, x1;
        SAF_FOR_P { acc.v[p] = 0.0; x0.v[p] = 0.0; x1.v[p] = 0.0; }

        for (unsigned long long it = 0;;) {
            // ---- the interpreter loop: warp-uniform pc / opcode ----
            uint32_t pc = 0;
            for (;;) {
                const uint64_t w = fetch_word(pc++);
                const uint32_t op = (uint32_t)w & 0xffu;
                if (op == D_END) break;
                const uint32_t fd = (uint32_t)(w >> 8) & FIELD_MASK;
                const uint32_t fa = (uint32_t)(w >> 22) & FIELD_MASK;
                const uint32_t fb = (uint32_t)(w >> 36) & FIELD_MASK;
                const uint32_t fc = (uint32_t)(w >> 50) & FIELD_MASK;

                auto operand = [&](uint32_t f, Pt<P> &o) {
                    const uint32_t kind = f >> INDEX_BITS, idx = f & INDEX_MASK;
                    if (kind == KIND_S) {
                        slot_load<P>(my, stride, idx, o);
                    } else if (kind == KIND_K) {
                        const double k = fetch_const(idx);
                        SAF_FOR_P o.v[p] = k;
                    } else {
                        o = acc;
                    }
                };

                Pt<P> a, b, c, r;
                operand(fa, a);
                if (op >= D_FIRST_BINARY) operand(fb, b);
                if (op >= D_FIRST_TERNARY) operand(fc, c);

                switch (op) {
                case D_COPY: r = a; break;
                case D_NEG: SAF_FOR_P r.v[p] = -a.v[p]; break;
                case D_SQUARE: SAF_FOR_P r.v[p] = __dmul_rn(a.v[p], a.v[p]); break;
                case D_CUBE: SAF_FOR_P r.v[p] = __dmul_rn(__dmul_rn(a.v[p], a.v[p]), a.v[p]); break;
                case D_POW4: SAF_FOR_P { double t = __dmul_rn(a.v[p], a.v[p]); r.v[p] = __dmul_rn(t, t); } break;
                case D_POW3_2: SAF_FOR_P r.v[p] = __dmul_rn(a.v[p], __dsqrt_rn(a.v[p])); break;
                case D_INVPOW3_2: SAF_FOR_P r.v[p] = __ddiv_rn(1.0, __dmul_rn(a.v[p], __dsqrt_rn(a.v[p]))); break;
                case D_INVSQRT: SAF_FOR_P r.v[p] = __ddiv_rn(1.0, __dsqrt_rn(a.v[p])); break;
                case D_INVSQUARE: SAF_FOR_P r.v[p] = __ddiv_rn(1.0, __dmul_rn(a.v[p], a.v[p])); break;
                case D_INVCUBE: SAF_FOR_P r.v[p] = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(a.v[p], a.v[p]), a.v[p])); break;
                case D_RECIP: SAF_FOR_P r.v[p] = __ddiv_rn(1.0, a.v[p]); break;
                case D_SIN: SAF_FOR_P r.v[p] = sin(a.v[p]); break;
                case D_COS: SAF_FOR_P r.v[p] = cos(a.v[p]); break;
                case D_EXP: SAF_FOR_P r.v[p] = exp(a.v[p]); break;
                case D_LN: SAF_FOR_P r.v[p] = log(a.v[p]); break;
                case D_SQRT: SAF_FOR_P r.v[p] = __dsqrt_rn(a.v[p]); break;
                case D_RECIPEXPM1: SAF_FOR_P r.v[p] = __ddiv_rn(1.0, expm1(a.v[p])); break;
                case D_EXPSQR: SAF_FOR_P r.v[p] = exp(__dmul_rn(a.v[p], a.v[p])); break;
                case D_EXPSQRNEG: SAF_FOR_P r.v[p] = exp(-__dmul_rn(a.v[p], a.v[p])); break;
                case D_EXPNEG: SAF_FOR_P r.v[p] = exp(-a.v[p]); break;
                case D_SINCOS: {
                    Pt<P> cs;
                    SAF_FOR_P sincos(a.v[p], &r.v[p], &cs.v[p]);
                    if ((fc >> INDEX_BITS) == KIND_S) slot_store<P>(my, stride, fc & INDEX_MASK, cs);
                } break;
                case D_BUILTIN1: {
                    const uint32_t fn = fb & INDEX_MASK;
#pragma unroll 1
                    for (int p = 0; p < P; ++p) r.v[p] = dm::builtin1(fn, a.v[p]);
                } break;
                case D_BUILTIN3: {
                    const uint32_t fn = fb & INDEX_MASK;
#pragma unroll 1
                    for (int p = 0; p < P; ++p) r.v[p] = dm::builtin3(fn, x0.v[p], x1.v[p], a.v[p]);
                } break;
                case D_ADD: SAF_FOR_P r.v[p] = __dadd_rn(a.v[p], b.v[p]); break;
                case D_SUB: SAF_FOR_P r.v[p] = __dsub_rn(a.v[p], b.v[p]); break;
                case D_MUL: SAF_FOR_P r.v[p] = __dmul_rn(a.v[p], b.v[p]); break;
                case D_DIV: SAF_FOR_P r.v[p] = __ddiv_rn(a.v[p], b.v[p]); break;
                case D_POW: SAF_FOR_P r.v[p] = pow(a.v[p], b.v[p]); break;
                case D_NEGMUL: SAF_FOR_P r.v[p] = -__dmul_rn(a.v[p], b.v[p]); break;
                case D_POWI: {
                    const int n = (int)b.v[0]; // K field, warp-uniform
                    SAF_FOR_P r.v[p] = dm::powi(a.v[p], n);
                } break;
                case D_BUILTIN2: {
                    const uint32_t fn = fc & INDEX_MASK;
#pragma unroll 1
                    for (int p = 0; p < P; ++p) r.v[p] = dm::builtin2(fn, a.v[p], b.v[p]);
                } break;
                case D_BUILTIN4: {
                    const uint32_t fn = fc & INDEX_MASK;
#pragma unroll 1
                    for (int p = 0; p < P; ++p) r.v[p] = dm::builtin4(fn, x0.v[p], x1.v[p], a.v[p], b.v[p]);
                } break;
                case D_ARG2:
                    x0 = a;
                    x1 = b;
                    continue; // stages operands only: ACC and slots untouched
                case D_FMA: SAF_FOR_P r.v[p] = __fma_rn(a.v[p], b.v[p], c.v[p]); break;
                case D_FMS: SAF_FOR_P r.v[p] = __fma_rn(a.v[p], b.v[p], -c.v[p]); break;
                case D_FNMA: SAF_FOR_P r.v[p] = __fma_rn(-a.v[p], b.v[p], c.v[p]); break;
                case D_FNMS: SAF_FOR_P r.v[p] = __fma_rn(-a.v[p], b.v[p], -c.v[p]); break;
                default: SAF_FOR_P r.v[p] = dm::nan_(); break;
                }
                acc = r;
                if ((fd >> INDEX_BITS) == KIND_S) slot_store<P>(my, stride, fd & INDEX_MASK, r);
            }

            if (++it >= d.iters) break;
            // ---- in-kernel feedback: parameter j <- result j (result slots never alias parameter
            // slots, lower.cpp), state stays on chip between iterations ----
            for (uint32_t j = 0; j < d.n_params; ++j) {
                const uint32_t s = d.param_slot[j];
                if (s == 0xffffu) continue;
                const uint32_t f = d.result_field[j];
                Pt<P> v;
                if ((f >> INDEX_BITS) == KIND_S) {
                    slot_load<P>(my, stride, f & INDEX_MASK, v);
                } else {
                    const double k = fetch_const(f & INDEX_MASK);
                    SAF_FOR_P v.v[p] = k;
                }
                slot_store<P>(my, stride, s, v);
            }
        }

        // ---- epilogue: result slots -> output columns ----
        for (uint32_t k = 0; k < d.n_results; ++k) {
            const uint32_t f = d.result_field[k];
            Pt<P> v;
            if ((f >> INDEX_BITS) == KIND_S) {
                slot_load<P>(my, stride, f & INDEX_MASK, v);
            } else {
                const double kv = fetch_const(f & INDEX_MASK);
                SAF_FOR_P v.v[p] = kv;
            }
            double *out = d.outs[k];
            if (P == 2 && i0 + 1 < d.n_points) {
                st_stream2(out + i0, v.v[0], v.v[P - 1]);
            } else {
                SAF_FOR_P if (i0 + p < d.n_points) st_stream(out + i0 + p, v.v[p]);
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

template <int P, int WORDS>
static cudaError_t launch_one(const LaunchDesc &desc, const uint64_t *words, size_t n_words,
                              const LaunchShape &shape, cudaStream_t stream) {
    auto kern = saf_interp_kernel<P, WORDS>;
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
    ConstBlob<(WORDS > 0 ? WORDS : 1)> blob;
    if (WORDS > 0) {
        for (size_t i = 0; i < n_words; ++i) blob.w[i] = words[i];
    }
    kern<<<shape.grid, shape.block, shape.smem_bytes, stream>>>(desc, blob);
    return cudaGetLastError();
}

template <int P>
static cudaError_t launch_p(const LaunchDesc &desc, const uint64_t *words, size_t n_words,
                            const LaunchShape &shape, cudaStream_t stream) {
    if (words == nullptr || n_words > (size_t)BLOB_LARGE) return launch_one<P, 0>(desc, nullptr, 0, shape, stream);
    if (n_words <= (size_t)BLOB_SMALL) return launch_one<P, BLOB_SMALL>(desc, words, n_words, shape, stream);
    if (n_words <= (size_t)BLOB_MEDIUM) return launch_one<P, BLOB_MEDIUM>(desc, words, n_words, shape, stream);
    return launch_one<P, BLOB_LARGE>(desc, words, n_words, shape, stream);
}

cudaError_t launch_interpreter(const LaunchDesc &desc, const uint64_t *blob_words, size_t n_blob_words,
                               const LaunchShape &shape, cudaStream_t stream) {
    if (shape.points_per_thread == 2) return launch_p<2>(desc, blob_words, n_blob_words, shape, stream);
    return launch_p<1>(desc, blob_words, n_blob_words, shape, stream);
}

cudaError_t choose_shape(int device, uint32_t n_slots, unsigned long long n_points, bool aligned16,
                         LaunchShape *shape) {
    int sms = 0, smem_optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    const size_t smem_sm = (size_t)smem_optin; // per-block opt-in limit ~= per-SM capacity minus 1 KB

    // Two points per thread once there is enough work to keep every SM busy with 2-point threads.
    int P = (aligned16 && n_points >= (unsigned long long)sms * 2048ull) ? 2 : 1;
    int block = 256;
    auto bytes = [&](int blk, int p) { return (size_t)n_slots * (size_t)blk * (size_t)p * 8; };
    // keep at least two resident blocks per SM when the register file of the program allows it
    while (block > 32 && bytes(block, P) * 2 > smem_sm) {
        if (P == 2 && block <= 128) P = 1; else block /= 2;
    }
    if (bytes(block, P) > smem_sm) {
        if (P == 2) P = 1;
        while (block > 32 && bytes(block, P) > smem_sm) block /= 2;
    }
    if (bytes(block, P) > smem_sm) return cudaErrorInvalidConfiguration;
    // small problems: shrink the block so that every SM gets work
    while (block > 64 && n_points < (unsigned long long)sms * (unsigned long long)block * (unsigned long long)P)
        block /= 2;

    const size_t smem = bytes(block, P);
    size_t blocks_by_smem = smem ? (smem_sm + 1024) / (smem + 1024) : 32;
    if (blocks_by_smem < 1) blocks_by_smem = 1;
    size_t blocks_by_threads = 2048 / (size_t)block;
    size_t resident = blocks_by_smem < blocks_by_threads ? blocks_by_smem : blocks_by_threads;
    if (resident > 32) resident = 32;
    const unsigned long long tile = (unsigned long long)block * P;
    unsigned long long n_tiles = (n_points + tile - 1) / tile;
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
