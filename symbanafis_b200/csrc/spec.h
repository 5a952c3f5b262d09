// spec.h -- specialised kernels: the device program of one saf_program written out as straight-line sm_100a code and
// compiled at run time (spec.cpp: text generation, jit.cpp: NVRTC, api.cu: saf_program_specialize / launches).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "lower.h"
#include "saf_isa.h"

namespace saf {

// Kernel parameter block of every specialised kernel (device mirror: spec_prelude.cuh sp::SpecDesc).
struct SpecDesc {
    unsigned long long n_points;
    unsigned long long iters;
    const double *cols[MAX_COLS];
    unsigned long long col_len[MAX_COLS];
    double *outs[MAX_OUTS];
};

struct SpecShape {
    int points_per_thread = 1; // 1, 2 or 4
    int block = 128;           // threads per block
    int min_blocks = 1;        // __launch_bounds__ second argument: caps the registers per thread
    int sync_every = 0;        // > 0: a block barrier every so many statements (long programs, spec.cpp)
    bool outline = false;      // transcendental leaves are called, not inlined
    bool prefetch = true;      // short programs: the next tile's columns are loaded before the current tile is evaluated
    bool speculate = true;
    bool rolled = false;       // the body is a loop over similar terms (reroll.inc); set by spec_generate
    std::string reroll_note;   // what was rolled, or why not     // inlined leaves: fast paths first, careful replay of the iteration when one did not apply
};

// A loaded specialised kernel on one device (launch.cu).
struct SpecModule {
    void *module = nullptr;   // CUmodule
    void *function = nullptr; // CUfunction `saf_spec`
    void *function_pf = nullptr; // CUfunction `saf_spec_pf` (single passes with next-tile prefetch) or null
    int registers = 0, local_bytes = 0, max_threads = 0, registers_pf = 0;
    int sms = 0, resident_blocks = 0, resident_blocks_pf = 0;
};

// Largest device program (instructions) that is specialised: beyond it compile time is better spent interpreting.
constexpr uint32_t SPEC_MAX_INSTRUCTIONS = 20000;

// Picks points per thread / block / launch bounds for a program.  Every program has two shapes: the one for batches
// that fill the machine, and (small_batch) one point per thread in narrow blocks for batches that do not -- there the
// warps, not the points per thread, are what hides latency.  Environment overrides (tuning knobs): SAF_SPEC_P,
// SAF_SPEC_BLOCK, SAF_SPEC_MINB, SAF_SPEC_SYNC, SAF_SPEC_INLINE_MAX.
SpecShape spec_shape_for(const DeviceProgram &dp, bool small_batch);
// Is a batch of n_points small for the large-batch shape `large` on a device with `sms` multiprocessors?
bool spec_batch_is_small(const SpecShape &large, unsigned long long n_points, int sms);

// Writes the CUDA C++ text of the kernel `saf_spec` for `dp`.  can_iterate: the text contains the in-kernel feedback
// of results into parameters (needs as many results as parameters).  Returns 0 or SAF_ERR_UNSUPPORTED.
// has_prefetch_entry: the text holds a second kernel, `saf_spec_pf`, for single passes (iters == 1): it loads the columns
// of a block's next tile before it evaluates the current one.
int spec_generate(const DeviceProgram &dp, SpecShape &shape, std::string &text, bool &can_iterate, bool &has_prefetch_entry,
                  std::string &err);

// NVRTC (jit.cpp).  Compiles `text` (which includes spec_prelude.cuh and through it the device math headers; all of
// them are embedded in this library) for sm_100a.  Returns 0, SAF_ERR_UNSUPPORTED when libnvrtc cannot be loaded, or
// SAF_ERR_CUDA with the compiler's log in err.
int jit_compile(const std::string &text, std::vector<char> &cubin, std::string &err);
// Is an NVRTC that can target sm_100a loadable in this process?
bool jit_available(std::string &why_not);


#ifdef __CUDACC__
} // namespace saf
#include <cuda_runtime.h>
namespace saf {
cudaError_t spec_module_load(const void *cubin, bool with_prefetch_entry, SpecModule *out); // current device
cudaError_t spec_module_occupancy(SpecModule *m, int block);                                // fills resident_blocks*
void spec_module_unload(SpecModule *m);
cudaError_t spec_launch(const SpecModule &m, bool prefetch_entry, const SpecDesc &desc, int grid, int block, cudaStream_t stream);
#endif

} // namespace saf
