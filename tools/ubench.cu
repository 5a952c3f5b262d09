// ubench.cu -- hardware ground truths that the interpreter's design rests on (one B200, one run, a few seconds).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu && build/ubench
// Prints one line per probe:
//   dfma_lat        cycles per dependent DFMA (one warp)
//   dfma_tput       SM-cycles per warp-DFMA at 4/8/16 warps per SM, full and half-active warps
//   lds_lat         cycles per dependent LDS.64 (pointer chase)
//   lds_tput        shared-memory bytes per clock per SM (LDS.128, 16 warps)
//   tmem_ld / st    tcgen05.ld/st 32x32b.x8 bytes per clock per SM at 4/8/16 warps; ld->use latency
//   brx             cycles per brx.idx dispatch through a small jump table (one warp / 8 warps)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__global__ void k_dfma_lat(double *sink, long long *cyc, int n) {
    double a = 1.0 + threadIdx.x * 1e-9; const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 32; ++k) a = __fma_rn(a, m, c);
    }
    long long t1 = clock64();
    if (a == 123.0) sink[0] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// independent chains: 8 per thread; active lanes = `lanes` per warp
__global__ void k_dfma_tput(double *sink, long long *cyc, int n, int lanes) {
    if ((threadIdx.x & 31) >= lanes) return;
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + threadIdx.x * 1e-9 + k;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = __fma_rn(a[k], m, c);
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.0) sink[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// DFMA whose addend / multiplier comes from the constant bank (what the lean sin / cos polynomials issue), and DFMA
// with three distinct register operands: is either slower than the register-register-accumulator form of the probe?
__device__ __constant__ double k_coef[8] = {1e-9, 2e-9, 3e-9, 4e-9, 5e-9, 6e-9, 7e-9, 8e-9};
__global__ void k_dfma_const(double *sink, long long *cyc, int n, int mode) {
    double a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = 1.0 + threadIdx.x * 1e-9 + k; b[k] = 1.0000001 + k * 1e-9; }
    const double m = 1.0000001;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (mode == 0) a[k] = __fma_rn(a[k], m, k_coef[k]);          // constant-bank addend
                else if (mode == 1) a[k] = __fma_rn(a[k], k_coef[k], m);      // constant-bank multiplier
                else a[k] = __fma_rn(a[k], b[k], b[(k + 1) & 7]);             // three distinct registers
            }
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.0) sink[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds_lat(long long *cyc, int n) {
    __shared__ unsigned long long chain[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) chain[i] = ((i + 17) & 255) * 8;
    __syncthreads();
    unsigned long long p = 0; uint32_t base = (uint32_t)__cvta_generic_to_shared(chain);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(p) : "r"(base + (uint32_t)p));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = (long long)p; }
}

__global__ void k_lds_tput(double *sink, long long *cyc, int n) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * 16;
    double s0 = 0, s1 = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double x, y;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(base + (uint32_t)((k & 3) * blockDim.x * 16)));
            s0 += x; s1 += y;
        }
    }
    long long t1 = clock64();
    if (s0 + s1 == 123.0) sink[0] = s0;
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- tensor memory as a per-thread indexed register file: tcgen05.ld / st 32x32b.x8 ----
__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: loads only, 4 back-to-back then one wait; mode 1: stores only; mode 2: ld -> wait -> dependent add -> st
// (round trip); mode 3: single dependent load chain (address depends on the loaded value): latency
__global__ void k_tmem(uint32_t *sink, long long *cyc, int n, int mode) {
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&tbase_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tbase_s;
    // warp w touches lanes 32 * (w % 4) ..; warps sharing a lane quadrant use disjoint column ranges
    const int n_warps = blockDim.x >> 5, per_quadrant = (n_warps + 3) / 4;
    const uint32_t cols_per_warp = 512u / per_quadrant;
    const uint32_t my = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(warp >> 2) * cols_per_warp;
    uint32_t r[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { r[k] = threadIdx.x + k; acc[k] = 0; }
    for (uint32_t c = 0; c < cols_per_warp; c += 8) tm_st8(my + c, r);
    tm_wait_st();
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) {
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            uint32_t a[8], b[8], c[8], d[8];
            const uint32_t o = (uint32_t)(i * 32) & (cols_per_warp - 1);
            tm_ld8(my + o, a); tm_ld8(my + ((o + 8) & (cols_per_warp - 1)), b);
            tm_ld8(my + ((o + 16) & (cols_per_warp - 1)), c); tm_ld8(my + ((o + 24) & (cols_per_warp - 1)), d);
            tm_wait_ld();
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += a[k] ^ b[k] ^ c[k] ^ d[k];
        }
    } else if (mode == 1) {
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const uint32_t o = (uint32_t)(i * 32) & (cols_per_warp - 1);
            tm_st8(my + o, r); tm_st8(my + ((o + 8) & (cols_per_warp - 1)), r);
            tm_st8(my + ((o + 16) & (cols_per_warp - 1)), r); tm_st8(my + ((o + 24) & (cols_per_warp - 1)), r);
            r[0] += i;
        }
        tm_wait_st();
    } else if (mode == 2) {
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const uint32_t o = (uint32_t)(i * 8) & (cols_per_warp - 1);
            tm_ld8(my + o, r);
            tm_wait_ld();
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] += 1;
            tm_st8(my + ((o + 8) & (cols_per_warp - 1)), r);
            tm_wait_st();
        }
    } else {
        uint32_t o = 0;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            tm_ld8(my + o, r);
            tm_wait_ld();
            o = (r[0] & 7u) * 8u; // value-dependent address: one load latency per iteration
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k] + r[k];
    if (s == 0x12345678u) sink[0] = s;
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase));
}

// indirect-branch dispatch: a loop that jumps through a 16-entry table to tiny blocks (one DADD each)
__global__ void k_brx(double *sink, long long *cyc, const int *prog, int len, int n) {
    extern __shared__ int code[];
    for (int i = threadIdx.x; i < len; i += blockDim.x) code[i] = prog[i];
    __syncthreads();
    double a = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < n; ++it) {
#pragma unroll 1
        for (int pc = 0; pc < len; ++pc) {
            switch (code[pc]) { // nvcc: compare tree + small tables; good enough for an order of magnitude
            case 0: a = __dadd_rn(a, 1.0); break;
            case 1: a = __dmul_rn(a, 1.0000001); break;
            case 2: a = __fma_rn(a, 1.0000001, 0.5); break;
            case 3: a = __dsub_rn(a, 0.25); break;
            case 4: a = __dadd_rn(a, 2.0); break;
            case 5: a = __dmul_rn(a, 0.9999999); break;
            case 6: a = __fma_rn(a, 0.9999999, 0.125); break;
            case 7: a = __dsub_rn(a, 0.75); break;
            default: a = -a; break;
            }
        }
    }
    long long t1 = clock64();
    if (a == 123.0) sink[0] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    double *sink; long long *cyc; long long h[64];
    CK(cudaMalloc(&sink, 1024)); CK(cudaMalloc(&cyc, 64 * sizeof(long long)));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);

    { const int n = 2000; k_dfma_lat<<<1, 32>>>(sink, cyc, n); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost)); printf("dfma_lat %.2f cycles per dependent DFMA\n", (double)h[0] / (n * 32.0)); }
    for (int lanes : {32, 16, 8}) for (int warps : {4, 8, 16, 32}) {
        const int n = 2000; k_dfma_tput<<<1, warps * 32>>>(sink, cyc, n, lanes); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        const double instr = (double)warps * n * 32.0; // warp-level DFMAs on the SM
        printf("dfma_tput lanes=%d warps=%d: %.3f SM-cycles per warp-DFMA (%.1f lanes/clk/SM)\n", lanes, warps, h[0] / instr, instr * lanes / h[0]);
    }
    for (int mode = 0; mode < 3; ++mode) for (int warps : {8, 16}) {
        const int n = 2000; k_dfma_const<<<1, warps * 32>>>(sink, cyc, n, mode); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        const double instr = (double)warps * n * 32.0;
        const char *names[] = {"constant-bank addend", "constant-bank multiplier", "three registers"};
        printf("dfma %s warps=%d: %.3f SM-cycles per warp-DFMA (%.1f lanes/clk/SM)\n", names[mode], warps, h[0] / instr, instr * 32 / h[0]);
    }
    { const int n = 2000; k_lds_lat<<<1, 32>>>(cyc, n); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost)); printf("lds_lat %.2f cycles per dependent LDS.64\n", (double)h[0] / (n * 16.0)); }
    for (int warps : {4, 8, 16, 32}) {
        const int n = 2000; k_lds_tput<<<1, warps * 32, warps * 32 * 16 * 4>>>(sink, cyc, n); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("lds_tput warps=%d: %.1f B/clk/SM (LDS.128)\n", warps, (double)warps * 32 * 16 * 8.0 * n / h[0]);
    }
    for (int mode = 0; mode < 4; ++mode) for (int warps : {4, 8, 16}) {
        if (mode == 3 && warps != 4) continue;
        const int n = 4000; k_tmem<<<1, warps * 32>>>((uint32_t *)sink, cyc, n, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tmem mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        const char *names[] = {"ld x8 (4 in flight)", "st x8 (4 in flight)", "ld->add->st round trip", "dependent ld chain"};
        const double per_iter_bytes = (mode <= 1 ? 4.0 : mode == 2 ? 2.0 : 1.0) * 32 * 32 * warps;
        printf("tmem %s warps=%d: %.1f B/clk/SM, %.1f cycles per iteration\n", names[mode], warps, per_iter_bytes * n / h[0], (double)h[0] / n);
    }
    {
        int hp[64]; for (int i = 0; i < 64; ++i) hp[i] = (i * 5 + 3) & 7;
        int *dp; CK(cudaMalloc(&dp, sizeof(hp))); CK(cudaMemcpy(dp, hp, sizeof(hp), cudaMemcpyHostToDevice));
        for (int warps : {1, 4, 8, 16}) {
            const int n = 2000; k_brx<<<1, warps * 32, 64 * sizeof(int)>>>(sink, cyc, dp, 64, n); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("switch-dispatch warps=%d: %.1f cycles per dispatched op (per warp)\n", warps, (double)h[0] / (n * 64.0));
        }
    }
    return 0;
}
