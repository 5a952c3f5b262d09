// launch.cu -- host side of the interpreter kernel: loads the cubin that the build produced from kernel.cu
// (nvcc -ptx -> tools/patch_brx.py -> ptxas, see Makefile) and launches its entry points.
//
// The cubin is embedded in this library (.incbin below) and loaded once per device through the driver API,
// whose entry points come from cudaGetDriverEntryPoint, so the library links against the CUDA runtime only
// and still loads on machines without a driver (program creation / lowering / export work there).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <cuda.h>
#include <cuda_runtime.h>

#include "kernel.cuh"
#include "spec.h"

// the patched, ptxas-compiled kernel image
asm(".section .rodata\n"
    ".global saf_kernel_cubin\n"
    ".global saf_kernel_cubin_end\n"
    ".balign 16\n"
    "saf_kernel_cubin:\n"
    ".incbin \"" SAF_CUBIN_PATH "\"\n"
    "saf_kernel_cubin_end:\n"
    ".byte 0\n"
    ".previous\n");
extern "C" const unsigned char saf_kernel_cubin[];
// the register machine's image (SAF_MACHINE=r), loaded on first use only
asm(".section .rodata\n"
    ".global saf_kernel_r_cubin\n"
    ".balign 16\n"
    "saf_kernel_r_cubin:\n"
    ".incbin \"" SAF_CUBIN_R_PATH "\"\n"
    ".byte 0\n"
    ".previous\n");
extern "C" const unsigned char saf_kernel_r_cubin[];

namespace saf {

namespace {

struct Driver {
    CUresult (*module_load_data)(CUmodule *, const void *) = nullptr;
    CUresult (*module_get_function)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*func_set_attribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*launch)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                       void **, void **) = nullptr;
    CUresult (*module_unload)(CUmodule) = nullptr;
    CUresult (*occupancy)(int *, CUfunction, int, size_t) = nullptr;
    CUresult (*func_get_attribute)(int *, CUfunction_attribute, CUfunction) = nullptr;
    bool ready = false;
} g_drv;

struct DeviceModule {
    CUmodule mod = nullptr;
    CUmodule mod_r = nullptr;   // register machine, loaded on first use
    CUfunction interp[15] = {}; // [13] [14] = streaming entry points P = 4 / 2; // [6] [7] = P = 8 (not shipped), [8] = register machine (from mod_r), [9] [10] = wide blocks P = 2, [11] [12] = wide blocks P = 4
    CUfunction peak = nullptr;
    CUfunction clifford_probe = nullptr;
    bool ready = false;
    int smem_optin = 0;
};
DeviceModule g_modules[64];
std::mutex g_module_mu;

cudaError_t as_cuda(CUresult r) {
    if (r == CUDA_SUCCESS) return cudaSuccess;
    if (r == CUDA_ERROR_OUT_OF_MEMORY) return cudaErrorMemoryAllocation;
    if (r == CUDA_ERROR_INVALID_VALUE) return cudaErrorInvalidValue;
    if (r == CUDA_ERROR_LAUNCH_OUT_OF_RESOURCES) return cudaErrorLaunchOutOfResources;
    if (r == CUDA_ERROR_NO_BINARY_FOR_GPU) return cudaErrorNoKernelImageForDevice;
    if (r == CUDA_ERROR_INVALID_IMAGE) return cudaErrorInvalidKernelImage;
    return cudaErrorUnknown;
}

template <typename F>
cudaError_t entry(const char *name, F *out) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess) return e;
    if (q != cudaDriverEntryPointSuccess || fn == nullptr) return cudaErrorSymbolNotFound;
    *out = reinterpret_cast<F>(fn);
    return cudaSuccess;
}

cudaError_t module_for_current_device(DeviceModule **out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(g_module_mu);
    DeviceModule &m = g_modules[dev];
    if (!m.ready) {
        e = cudaFree(nullptr); // make sure the runtime's primary context exists and is current
        if (e != cudaSuccess) return e;
        if (!g_drv.ready) {
            if ((e = entry("cuModuleLoadData", &g_drv.module_load_data)) != cudaSuccess) return e;
            if ((e = entry("cuModuleGetFunction", &g_drv.module_get_function)) != cudaSuccess) return e;
            if ((e = entry("cuFuncSetAttribute", &g_drv.func_set_attribute)) != cudaSuccess) return e;
            if ((e = entry("cuLaunchKernel", &g_drv.launch)) != cudaSuccess) return e;
            if ((e = entry("cuModuleUnload", &g_drv.module_unload)) != cudaSuccess) return e;
            if ((e = entry("cuOccupancyMaxActiveBlocksPerMultiprocessor", &g_drv.occupancy)) != cudaSuccess) return e;
            if ((e = entry("cuFuncGetAttribute", &g_drv.func_get_attribute)) != cudaSuccess) return e;
            g_drv.ready = true;
        }
        CUresult r = g_drv.module_load_data(&m.mod, saf_kernel_cubin);
        if (r != CUDA_SUCCESS) return as_cuda(r);
        static const char *names[15] = {"saf_interp_p4_ksm", "saf_interp_p4_kgm", "saf_interp_p2_ksm", "saf_interp_p2_kgm",
                                       "saf_interp_p1_ksm", "saf_interp_p1_kgm", "saf_interp_p8_ksm", "saf_interp_p8_kgm",
                                       "saf_rinterp_ksm", "saf_interp_p2w_ksm", "saf_interp_p2w_kgm",
                                       "saf_interp_p4w_ksm", "saf_interp_p4w_kgm", "saf_interp_p4s_ksm", "saf_interp_p2s_ksm"};
        int optin = 0;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e != cudaSuccess) return e;
        m.smem_optin = optin;
        for (int k = 0; k < 15; ++k) {
            if (k == 8) continue; // register machine: module_r_for()
            r = g_drv.module_get_function(&m.interp[k], m.mod, names[k]);
            if (r == CUDA_ERROR_NOT_FOUND && (k == 6 || k == 7)) { // built without -DSAF_WITH_P8
                m.interp[k] = nullptr;
                continue;
            }
            if (r != CUDA_SUCCESS) return as_cuda(r);
            // opt in to the full dynamic shared-memory carve-out
            r = g_drv.func_set_attribute(m.interp[k], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, optin);
            if (r != CUDA_SUCCESS) return as_cuda(r);
        }
        r = g_drv.module_get_function(&m.peak, m.mod, "saf_fp64_peak_kernel");
        if (r != CUDA_SUCCESS) return as_cuda(r);
        r = g_drv.module_get_function(&m.clifford_probe, m.mod, "saf_clifford_probe_kernel");
        if (r != CUDA_SUCCESS) return as_cuda(r);
        m.ready = true;
    }
    *out = &m;
    return cudaSuccess;
}

// loads the register machine's cubin into the current device's module set (under g_module_mu)
cudaError_t module_r_for(DeviceModule *m) {
    std::lock_guard<std::mutex> lk(g_module_mu);
    if (m->interp[8]) return cudaSuccess;
    CUresult r = g_drv.module_load_data(&m->mod_r, saf_kernel_r_cubin);
    if (r != CUDA_SUCCESS) return as_cuda(r);
    CUfunction f = nullptr;
    r = g_drv.module_get_function(&f, m->mod_r, "saf_rinterp_ksm");
    if (r != CUDA_SUCCESS) return as_cuda(r);
    r = g_drv.func_set_attribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, m->smem_optin);
    if (r != CUDA_SUCCESS) return as_cuda(r);
    m->interp[8] = f;
    return cudaSuccess;
}

} // namespace


cudaError_t launch_interpreter(const LaunchDesc &desc, const LaunchShape &shape, cudaStream_t stream) {
    const int li = shape.points_per_thread == 4 ? 0 : shape.points_per_thread == 2 ? 1 : shape.points_per_thread == 1 ? 2 : 3;
    const bool wide = shape.block != BLOCK_THREADS;
    if (wide && ((shape.points_per_thread != 2 && shape.points_per_thread != 4) || shape.rmachine)) return cudaErrorInvalidConfiguration;
    const bool streaming = shape.n_rf > 1;
    if (streaming && (wide || shape.rmachine || !shape.k_in_smem || (shape.points_per_thread != 4 && shape.points_per_thread != 2)))
        return cudaErrorInvalidConfiguration;
    const int k = shape.rmachine ? 8 : streaming ? (shape.points_per_thread == 4 ? 13 : 14)
                  : wide ? (shape.points_per_thread == 4 ? 11 : 9) + (shape.k_in_smem ? 0 : 1) : 2 * li + (shape.k_in_smem ? 0 : 1);
    if (shape.rmachine && !shape.k_in_smem) return cudaErrorInvalidConfiguration;
    DeviceModule *m = nullptr;
    cudaError_t e = module_for_current_device(&m);
    if (e != cudaSuccess) return e;
    if (shape.rmachine && (e = module_r_for(m)) != cudaSuccess) return e;
    if (!m->interp[k]) return cudaErrorInvalidConfiguration; // layout not in this build
    void *args[] = {const_cast<LaunchDesc *>(&desc)};
    return as_cuda(g_drv.launch(m->interp[k], (unsigned)shape.grid, 1, 1, (unsigned)shape.block, 1, 1,
                                (unsigned)shape.smem_bytes, (CUstream)stream, args, nullptr));
}

// ---- specialised kernels (spec.h): one module per program and device, loaded from the cubin NVRTC produced ----------
cudaError_t spec_module_load(const void *cubin, bool with_prefetch_entry, SpecModule *out) {
    DeviceModule *m = nullptr;
    cudaError_t e = module_for_current_device(&m); // driver entry points + primary context
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    CUmodule mod = nullptr;
    CUresult r = g_drv.module_load_data(&mod, cubin);
    if (r != CUDA_SUCCESS) return as_cuda(r);
    CUfunction f = nullptr;
    r = g_drv.module_get_function(&f, mod, "saf_spec");
    int threads = 0, regs = 0, local = 0, resident = 0;
    if (r == CUDA_SUCCESS) r = g_drv.func_get_attribute(&threads, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, f);
    if (r == CUDA_SUCCESS) r = g_drv.func_get_attribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, f);
    if (r == CUDA_SUCCESS) r = g_drv.func_get_attribute(&local, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f);
    CUfunction fp = nullptr;
    int regs_pf = 0;
    if (r == CUDA_SUCCESS && with_prefetch_entry) {
        r = g_drv.module_get_function(&fp, mod, "saf_spec_pf");
        if (r == CUDA_SUCCESS) r = g_drv.func_get_attribute(&regs_pf, CU_FUNC_ATTRIBUTE_NUM_REGS, fp);
    }
    if (r != CUDA_SUCCESS) {
        g_drv.module_unload(mod);
        return as_cuda(r);
    }
    out->module = mod;
    out->function = f;
    out->function_pf = fp;
    out->registers_pf = regs_pf;
    out->resident_blocks_pf = 0;
    out->registers = regs;
    out->local_bytes = local;
    out->max_threads = threads;
    out->sms = sms;
    out->resident_blocks = resident;
    return cudaSuccess;
}

cudaError_t spec_module_occupancy(SpecModule *m, int block) {
    int resident = 0;
    CUresult r = g_drv.occupancy(&resident, (CUfunction)m->function, block, 0);
    if (r != CUDA_SUCCESS) return as_cuda(r);
    m->resident_blocks = resident < 1 ? 1 : resident;
    if (m->function_pf) {
        r = g_drv.occupancy(&resident, (CUfunction)m->function_pf, block, 0);
        if (r != CUDA_SUCCESS) return as_cuda(r);
        m->resident_blocks_pf = resident < 1 ? 1 : resident;
    }
    return cudaSuccess;
}

void spec_module_unload(SpecModule *m) {
    if (m->module && g_drv.module_unload) g_drv.module_unload((CUmodule)m->module);
    m->module = m->function = m->function_pf = nullptr;
}

cudaError_t spec_launch(const SpecModule &m, bool prefetch_entry, const SpecDesc &desc, int grid, int block, cudaStream_t stream) {
    void *args[] = {const_cast<SpecDesc *>(&desc)};
    return as_cuda(g_drv.launch((CUfunction)(prefetch_entry ? m.function_pf : m.function), (unsigned)grid, 1, 1, (unsigned)block, 1, 1, 0, (CUstream)stream, args, nullptr));
}

cudaError_t launch_fp64_peak(double *sink, int blocks, int iters, cudaStream_t stream) {
    DeviceModule *m = nullptr;
    cudaError_t e = module_for_current_device(&m);
    if (e != cudaSuccess) return e;
    double seed = 0.5;
    void *args[] = {&sink, &iters, &seed};
    return as_cuda(g_drv.launch(m->peak, (unsigned)blocks, 1, 1, 256, 1, 1, 0, (CUstream)stream, args, nullptr));
}

cudaError_t launch_clifford_probe(double *xs, double *ys, unsigned long long n, int iters, int blocks,
                                  cudaStream_t stream) {
    DeviceModule *m = nullptr;
    cudaError_t e = module_for_current_device(&m);
    if (e != cudaSuccess) return e;
    double a = -1.4, b = 1.6, c = 1.0, d = 0.7;
    void *args[] = {&xs, &ys, &n, &iters, &a, &b, &c, &d};
    return as_cuda(g_drv.launch(m->clifford_probe, (unsigned)blocks, 1, 1, BLOCK_THREADS, 1, 1, 0, (CUstream)stream, args,
                                nullptr));
}

// register file + block context + constant table (when staged) + code window
size_t interpreter_smem_bytes(uint32_t n_slots, int points_per_thread, size_t k_bytes_in_smem, size_t code_bytes,
                              int block_threads, int n_rf) {
    const size_t window = code_bytes < CODE_WINDOW_BYTES ? code_bytes : CODE_WINDOW_BYTES;
    return (size_t)n_rf * (size_t)n_slots * (size_t)block_threads * (size_t)points_per_thread * 8 + sizeof(BlockCtx) +
           ((k_bytes_in_smem + 15) & ~(size_t)15) + ((window + 15) & ~(size_t)15) + (n_rf > 1 ? STREAM_MBAR_BYTES : 0);
}

static int resident_blocks_for(int P) { return P == 8 ? SAF_MINB_P8 : P == 4 ? SAF_MINB_P4 : P == 2 ? SAF_MINB_P2 : SAF_MINB_P1; }

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
        return (interpreter_smem_bytes(n_slots + 3, p, 0, CODE_WINDOW_BYTES) + 1024) * (size_t)min_blocks <= smem_sm + 1024;
    };
    auto enough = [&](int p) {
        return n_points >= (unsigned long long)sms * 3ull * (unsigned long long)BLOCK_THREADS * (unsigned long long)p;
    };
    int P = 1;
    if (aligned16 && enough(4) && fits(4, 2)) P = 4;
    else if (aligned16 && enough(2) && fits(2, 2)) P = 2;
    if (force_points_per_thread == 1 ||
        (aligned16 && (force_points_per_thread == 2 || force_points_per_thread == 4 || force_points_per_thread == 8)))
        P = force_points_per_thread;
#ifndef SAF_WITH_P8
    if (P == 8) P = 4; // the eight-points layout is not in the shipped image
#endif
    while (P > 1 && !fits(P, 1)) P >>= 1;
    if (!fits(P, 1)) return cudaErrorInvalidConfiguration;
    *points_per_thread = P;
    return cudaSuccess;
}

int wide_block_threads(int device, uint32_t n_slots, size_t k_bytes, size_t code_bytes, unsigned long long n_points, int P) {
    int sms = 0, smem_optin = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
    const size_t k_smem = k_bytes <= K_SMEM_MAX ? k_bytes : 0;
    // warps per SM with 128-thread blocks
    const size_t small = interpreter_smem_bytes(n_slots, P, k_smem, code_bytes);
    size_t blocks = ((size_t)smem_optin + 1024) / (small + 1024);
    if (blocks > (size_t)resident_blocks_for(P)) blocks = resident_blocks_for(P);
    const int warps_small = (int)blocks * (BLOCK_THREADS / 32);
    int threads = 0;
    for (int t = WIDE_MAX_THREADS; t > BLOCK_THREADS; t -= 32)
        if (interpreter_smem_bytes(n_slots, P, k_smem, code_bytes, t) <= (size_t)smem_optin) {
            threads = t;
            break;
        }
    if (threads / 32 <= warps_small) return 0;
    // every SM must still get whole tiles to work on
    if (n_points < (unsigned long long)sms * 2ull * (unsigned long long)threads * (unsigned long long)P) return 0;
    return threads;
}

cudaError_t choose_shape(int device, uint32_t n_slots, size_t k_bytes, size_t code_bytes, unsigned long long n_points,
                         int P, bool rmachine, LaunchShape *shape, int block_threads, bool stream_ok) {
    int sms = 0, smem_optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    bool k_in_smem = k_bytes <= K_SMEM_MAX;
    size_t smem = interpreter_smem_bytes(n_slots, P, k_in_smem ? k_bytes : 0, code_bytes, block_threads);
    if (smem > (size_t)smem_optin && k_in_smem) {
        k_in_smem = false;
        smem = interpreter_smem_bytes(n_slots, P, 0, code_bytes, block_threads);
    }
    if (smem > (size_t)smem_optin) return cudaErrorInvalidConfiguration;
    size_t resident = rmachine ? (size_t)SAF_MINB_R : (size_t)resident_blocks_for(P); // __launch_bounds__ register budget
    const size_t by_smem = ((size_t)smem_optin + 1024) / (smem + 1024);
    if (by_smem < resident) resident = by_smem;
    if (resident < 1 || block_threads != BLOCK_THREADS) resident = 1;
    // Streaming: extra register-file copies that column tiles are bulk-copied into while another copy is being
    // evaluated.  Taken when they cost no resident block (HBM-bound programs are small: few slots).
    int n_rf = 1;
    if (stream_ok && !rmachine && (P == 2 || P == 4) && code_bytes <= CODE_WINDOW_BYTES && k_in_smem && block_threads == BLOCK_THREADS) {
        static const int want = []() {
            const char *v = getenv("SAF_STREAM_BUFS"); // tuning knob: 1 disables streaming, 2 (default) / 3 copies
            const int n = v ? atoi(v) : 2;
            return n < 1 ? 1 : n > 3 ? 3 : n;
        }();
        for (int r = want; r >= 2; --r) {
            const size_t sm = interpreter_smem_bytes(n_slots, P, k_in_smem ? k_bytes : 0, code_bytes, block_threads, r);
            if (sm <= (size_t)smem_optin && ((size_t)smem_optin + 1024) / (sm + 1024) >= resident) {
                n_rf = r;
                smem = sm;
                break;
            }
        }
    }
    const unsigned long long tile = (unsigned long long)block_threads * P;
    const unsigned long long n_tiles = (n_points + tile - 1) / tile;
    unsigned long long grid = (unsigned long long)sms * resident;
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    // Balance: with `grid` resident blocks the tiles take ceil(n_tiles / grid) rounds; the same number of rounds with
    // every block doing the SAME number of tiles needs only ceil(n_tiles / rounds) blocks -- no nearly empty last round
    // in which a few blocks hold the kernel while most of the machine idles (977 tiles on 740 slots: 489 x 2 instead
    // of 740 + 237).  Measured: Aizawa +2 %, Clifford -10 % (an FP64-bound kernel wants every resident warp in the
    // rounds that are full) -- off unless SAF_BALANCE=1 (tuning knob).
    static const bool balance = []() {
        const char *v = getenv("SAF_BALANCE");
        return v && v[0] == '1';
    }();
    if (balance && n_tiles > grid) {
        const unsigned long long rounds = (n_tiles + grid - 1) / grid;
        grid = (n_tiles + rounds - 1) / rounds;
    }
    shape->n_rf = n_rf;
    shape->points_per_thread = P;
    shape->block = block_threads;
    shape->grid = (int)grid;
    shape->smem_bytes = smem;
    shape->k_in_smem = k_in_smem;
    shape->rmachine = rmachine;
    return cudaSuccess;
}

} // namespace saf
