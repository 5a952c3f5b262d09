// lower.h -- host-side lowering of reference bytecode to the device program (see saf_isa.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace saf {

// The reference's program payload (fields of CompiledEvaluator, src/evaluator/api.rs:125-142).
struct RefProgram {
    std::vector<uint32_t> flat;
    std::vector<double> consts;
    std::vector<uint32_t> pool;
    uint32_t param_count = 0;
    uint32_t workspace_size = 0;
    std::vector<uint32_t> result_regs;
};

struct DeviceProgram {
    std::vector<uint64_t> code;         // device word stream (lo [+ hi] per instruction), X_END terminated
    std::vector<double> consts;         // K table
    uint32_t n_slots = 0;               // S slots per point
    uint32_t param_count = 0;
    std::vector<uint16_t> param_slot;   // slot of each parameter, 0xffff when the program never reads it
    std::vector<uint32_t> result_field; // S or K field of each result
    // statistics
    uint32_t n_ref_instructions = 0;
    uint32_t n_dev_instructions = 0;    // excluding the end marker
    uint32_t n_acc_forwards = 0;
    uint32_t n_prefix_fusions = 0;
    uint32_t quirk_flags = 0;
    double fp64_slots = 0.0;
};

constexpr uint16_t NO_SLOT = 0xffff;

// Always-on validator: bounds, read-before-definition, arg-pool ranges, stream well-formedness
// (semantics of compile/optimize/helper.rs:98-196, which the reference only runs in debug builds).
// Returns 0 or SAF_ERR_INVALID_PROGRAM and fills err.
int validate(const RefProgram &p, std::string &err);

// Lowers one or several programs over the same parameter list into one device program.
// Returns 0, SAF_ERR_INVALID_PROGRAM or SAF_ERR_UNSUPPORTED.
int lower(const std::vector<const RefProgram *> &progs, DeviceProgram &out, std::string &err);

std::string disassemble(const DeviceProgram &p);

} // namespace saf
