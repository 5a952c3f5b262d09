// kernel.cuh -- launch descriptor shared by the host API (api.cu) and the interpreter kernel (kernel.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "saf_isa.h"
#include "uop.h"

// resident blocks per SM the register allocation aims for, per points-per-thread variant (tuning knobs)
#ifndef SAF_MINB_P4
#define SAF_MINB_P4 6
#endif
#ifndef SAF_MINB_P8
#define SAF_MINB_P8 4
#endif
#ifndef SAF_MINB_P2
#define SAF_MINB_P2 6
#endif
#ifndef SAF_MINB_R
#define SAF_MINB_R 4
#endif
#ifndef SAF_MINB_P1
#define SAF_MINB_P1 6
#endif


namespace saf {

// Everything the kernel needs besides the program image, passed by value in the kernel parameter
// (constant) bank.  Offsets are BYTES: register-file offsets are relative to the block's shared
// memory, image offsets relative to the start of the packed image (pack.h).
struct LaunchDesc {
    unsigned long long n_points;
    unsigned long long iters;            // >= 1; > 1 only for saf_iterate_*
    const unsigned char *image_global;   // packed image (pack.h) in global memory
    uint32_t code_off;                   // byte offset of the code in the image = size of the constant table
    uint32_t code_bytes;
    uint32_t n_slots;                    // register-file slots (incl. staging slots)
    uint32_t n_params;                   // columns the program can read (param_count)
    uint32_t n_feedback;                 // parameters copied back from results between iterations (0 = double buffered)
    uint32_t n_results;
    uint32_t flags;                      // SAF_EVAL_BROADCAST
    uint32_t n_rf;                       // register-file copies: 1, or 2 / 3 = STREAMING (bulk-copied column tiles, below)
    uint32_t result_k_mask;              // bit k: result k is a constant (result_off = image offset)
    int32_t acc_off;                     // reserved slot through which ACC is handed to the cold micro-ops
    int32_t x0_off, x1_off;              // staging slots of Builtin3/4 operands
    const double *cols[MAX_COLS];        // device column base pointers (nullptr = missing -> 0.0)
    unsigned long long col_len[MAX_COLS];
    double *outs[MAX_OUTS];
    int32_t param_off[MAX_COLS];         // register-file offset of each parameter (-1 = never read)
    int32_t result_off[MAX_OUTS];
};

// Constant tables up to this size are staged in shared memory (larger ones are read from global memory).
constexpr size_t K_SMEM_MAX = 24 * 1024;

struct LaunchShape {
    int n_rf = 1;          // register-file copies in shared memory (> 1: streaming single passes, kernel.cu)
    int points_per_thread; // 1, 2, 4 or 8
    int block;             // threads per block: BLOCK_THREADS, or more for the wide one-block-per-SM layout (P = 2)
    int grid;
    size_t smem_bytes;
    bool k_in_smem;        // the constant table is staged in shared memory (else read from global memory)
    bool rmachine;         // register machine (ruop.h): saf_rinterp_ksm
};

// Launches the interpreter; desc.image_global must point at a device copy of the packed image.
cudaError_t launch_interpreter(const LaunchDesc &desc, const LaunchShape &shape, cudaStream_t stream);

// Points per thread for a program with n_slots live values (before reserved slots) on `device`.
// force_points_per_thread: 0 = automatic, 1 / 2 / 4 / 8 = tuning override (SAF_POINTS_PER_THREAD).
cudaError_t choose_points_per_thread(int device, uint32_t n_slots, unsigned long long n_points, bool aligned16,
                                     int force_points_per_thread, int *points_per_thread);
// Grid / shared memory / constant-table placement for the packed program.
// stream_ok: the launch qualifies for streaming (single pass over full-length aligned columns, see kernel.cu); the shape
// then carries n_rf = 2 or 3 when the extra register-file copies cost no resident block.
cudaError_t choose_shape(int device, uint32_t n_slots, size_t k_bytes, size_t code_bytes, unsigned long long n_points,
                         int points_per_thread, bool rmachine, LaunchShape *shape, int block_threads = BLOCK_THREADS,
                         bool stream_ok = false);
// n_rf register files + block context + constant table (when staged) + code window
size_t interpreter_smem_bytes(uint32_t n_slots, int points_per_thread, size_t k_bytes_in_smem, size_t code_bytes,
                              int block_threads = BLOCK_THREADS, int n_rf = 1);
// Threads of the wide one-block-per-SM layout (two points per thread) for a packed program with n_slots slots, or 0
// when it would not hold more warps per SM than 128-thread blocks do.
int wide_block_threads(int device, uint32_t n_slots, size_t k_bytes, size_t code_bytes, unsigned long long n_points,
                       int points_per_thread = 2);

// Per-block context behind the register file in shared memory: what hot_run needs besides the image.
struct alignas(16) BlockCtx { // a multiple of 16 bytes: the constant table and the code window follow it
    uint32_t n_params;
    uint32_t result_k_mask;
    int32_t acc_off;
    uint32_t code_off;
    uint32_t code_bytes;
    uint32_t pad_;
    int32_t param_off[MAX_COLS];
    int32_t result_off[MAX_OUTS];
};
// Streaming launches (n_rf > 1) append one mbarrier per register-file copy behind the code window ("the column tiles
// of that copy have landed"): STREAM_MBAR_BYTES, counted by interpreter_smem_bytes.
constexpr uint32_t STREAM_MAX_COPIES = 3;
constexpr uint32_t STREAM_MBAR_BYTES = 32;

static_assert(sizeof(BlockCtx) % 16 == 0, "the code window behind BlockCtx is fetched with 16-byte loads");

// FP64-pipe peak probe: blocks x 256 threads, each `iters` x 64 dependent-chain DFMAs.
cudaError_t launch_fp64_peak(double *sink, int blocks, int iters, cudaStream_t stream);
// Hard-coded Clifford map with the same lean sin/cos kernels (no interpreter): the "ideal kernel" reference point.
cudaError_t launch_clifford_probe(double *xs, double *ys, unsigned long long n, int iters, int blocks,
                                  cudaStream_t stream);

} // namespace saf
