// kernel.cuh -- launch descriptor shared by the host API (api.cu) and the interpreter kernel (kernel.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "saf_isa.h"

namespace saf {

// Everything the kernel needs besides the program words, passed by value in the kernel parameter
// (constant) bank.
struct LaunchDesc {
    unsigned long long n_points;
    unsigned long long iters;            // >= 1; > 1 only for saf_iterate_*
    const uint64_t *code_global;         // device words in global memory (used when the program does
    const double *consts_global;         //   not fit the constant-bank blob)
    uint32_t n_code;                     // device words including D_END
    uint32_t n_consts;
    uint32_t n_slots;
    uint32_t n_params;                   // columns the program can read (param_count)
    uint32_t n_results;
    uint32_t flags;                      // SAF_EVAL_BROADCAST
    const double *cols[MAX_COLS];        // device column base pointers (nullptr = missing -> 0.0)
    unsigned long long col_len[MAX_COLS];
    double *outs[MAX_OUTS];
    uint16_t param_slot[MAX_COLS];       // slot of each parameter (0xffff = never read)
    uint16_t result_field[MAX_OUTS];     // S / K field of each result
};

// Program words + constants staged in the kernel parameter space, i.e. the constant bank: fetched
// through the constant cache with warp-uniform addresses, off the shared-memory / L1 data path.
template <int WORDS>
struct ConstBlob {
    uint64_t w[WORDS]; // [0, n_code) device words, then n_consts f64 bit patterns
};

// Blob size classes (8-byte words).  With LaunchDesc (< 1 KB) the largest stays under the 32764-byte
// kernel parameter limit of CUDA 12.1+.
constexpr int BLOB_SMALL = 448;
constexpr int BLOB_LARGE = 3840;

struct LaunchShape {
    int points_per_thread; // 1, 2 or 4
    int block;             // threads per block (always BLOCK_THREADS)
    int grid;
    size_t smem_bytes;
};

// Launches the interpreter.  `blob_words` holds the code words (followed by the constants when
// `consts_in_blob`) if they fit BLOB_LARGE; otherwise pass nullptr and the kernel reads
// desc.code_global / desc.consts_global.
cudaError_t launch_interpreter(const LaunchDesc &desc, const uint64_t *blob_words, size_t n_blob_words,
                               bool consts_in_blob, const LaunchShape &shape, cudaStream_t stream);

// Picks points-per-thread / grid for a program with n_slots live values on `device`.
// force_points_per_thread: 0 = automatic, 1 / 2 / 4 = tuning override (SAF_POINTS_PER_THREAD).
cudaError_t choose_shape(int device, uint32_t n_slots, unsigned long long n_points, bool aligned16,
                         int force_points_per_thread, LaunchShape *shape);
size_t interpreter_smem_bytes(uint32_t n_slots, int points_per_thread);

// FP64-pipe peak probe: blocks x 256 threads, each `iters` x 64 dependent-chain DFMAs.
cudaError_t launch_fp64_peak(double *sink, int blocks, int iters, cudaStream_t stream);

} // namespace saf
