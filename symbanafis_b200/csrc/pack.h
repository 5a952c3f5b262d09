// pack.h -- packs a device program (saf_isa.h) into the kernel's micro-instruction image (uop.h) for
// one register-file layout.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "lower.h"
#include "uop.h"

namespace saf {

struct PackedProgram {
    int points_per_thread = 0;
    int block_threads = BLOCK_THREADS; // threads per block the slot stride was computed for
    int layout_id = 0;                 // points_per_thread (128-thread blocks), 1000 * P + threads (wide blocks), 100 + P (register machine)
    bool can_iterate = true;           // the image loops in-kernel (saf_iterate_*); register-machine images of programs
                                       // whose results do not pair with their parameters cannot
    // image = constant table (padded to 32 bytes), then the code: n_instr micro-instruction slots cut into
    // windows of CODE_WINDOW_BYTES (uop.h).  K operands are byte offsets into the image, program counters are
    // byte offsets into the code.
    std::vector<uint64_t> image;
    uint32_t code_off = 0;             // byte offset of the code in the image = size of the constant table
    uint32_t code_bytes = 0;
    uint32_t n_instr = 0;
    uint32_t n_slots = 0;              // register-file slots incl. the reserved ones below
    int32_t acc_off = 0;               // reserved slot through which ACC is handed to cold micro-ops / heavy ops
    int32_t x0_off = -1, x1_off = -1;  // staging slots of Builtin3/4 operands (byte offsets), -1 when unused
    std::vector<int32_t> param_off;    // per parameter: register-file byte offset, -1 = never read
    std::vector<int32_t> result_off;   // per result: byte offset (register file or image)
    // Iterated maps: the image holds the program twice with parameter and result slots swapped in the second
    // copy (no copies between iterations); after an EVEN number of iterations results sit at result_off_even.
    bool double_buffered = false;
    std::vector<int32_t> result_off_even;
    std::vector<uint8_t> result_is_k;  // per result: 1 = constant (offset into the image)
};

// Byte stride of one slot for P points per thread.
constexpr uint32_t slot_bytes(int points_per_thread, int block_threads = BLOCK_THREADS) {
    return (uint32_t)block_threads * 8u * (uint32_t)points_per_thread;
}

// Returns 0 or SAF_ERR_UNSUPPORTED (err filled).
// Register machine (ruop.h, pack_r.cpp): four points per thread, live values in real registers.
int pack_program_r(const DeviceProgram &dp, PackedProgram &out, std::string &err);
// block_threads != BLOCK_THREADS: the wide one-block-per-SM layout (two or four points per thread), layout_id = 1000 * P + threads.
int pack_program(const DeviceProgram &dp, int points_per_thread, PackedProgram &out, std::string &err,
                 int block_threads = BLOCK_THREADS);

} // namespace saf
