// api.cu -- implementation of the C ABI declared in include/saf_b200.h.
//
// Host-side mirror of the reference drivers:
//   run_chunked_evaluator   src/evaluator/logic/bytecode/execute/drivers/batch.rs:37-78
//   eval_batch              .../execute/engine/scalar.rs:150-230
// The Rayon par_chunks_mut(256) of the reference becomes: one interpreter launch per chunk of points,
// chunks pipelined over internal streams (H2D copy / kernel / D2H copy overlap) for host buffers,
// one launch for device buffers.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <unistd.h>

#include "../../include/saf_b200.h"
#include "kernel.cuh"
#include "lower.h"
#include "pack.h"
#include "ruop.h"
#include "spec.h"

using namespace saf;

// ---- error plumbing ---------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char *what) {
    std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return fail(SAF_ERR_CUDA, m);
}
#define SAF_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);                                        \
    } while (0)

// ---- program handle ---------------------------------------------------------------------------------
struct saf_program {
    std::vector<RefProgram> refs; // kept so that handles can be fused later
    DeviceProgram dev;
    std::mutex mu;
    // packed micro-instruction images, one per register-file layout (P = 1, 2, 4), built on first use
    PackedProgram packed[5];          // [4] = register machine (ruop.h)
    bool packed_ready[5] = {false, false, false, false, false};
    int r_pack_rc = 0;                // result of the register-machine packing (kept: failures are not retried)
    // wide one-block-per-SM layout, keyed by 1000 * points-per-thread + block threads.  std::map never moves its
    // entries: a pointer handed out under `mu` stays valid after the lock is released (devices with different
    // shared-memory sizes may ask for different widths concurrently, saf_eval_host_multi)
    std::map<int, PackedProgram> packed_wide;
    std::map<std::pair<int, int>, unsigned char *> copies; // (device, P) -> global-memory image for big programs
    // specialised kernel (spec.h): generated and compiled once by saf_program_specialize, loaded per device on first use
    // [0]: the shape for batches that fill the machine (compiled by saf_program_specialize); [1]: the small-batch shape,
    // compiled the first time a launch needs it (spec.h spec_batch_is_small)
    struct Spec {
        bool attempted = false;
        int rc = SAF_OK;
        std::string err, text;
        std::vector<char> cubin;
        SpecShape shape;
        bool can_iterate = false, has_prefetch_entry = false;
        double generate_ms = 0.0, compile_ms = 0.0;
        std::map<int, SpecModule> modules; // by device; entries never move
    } spec[2];
    double interpreted_work = 0.0; // SAF_SPECIALIZE=auto: instruction-points run on the interpreter so far
};

constexpr int LAYOUT_R = 100 + R_POINTS; // layout id of the register machine
static int layout_index(int P) { return P == LAYOUT_R ? 4 : P == 8 ? 3 : P == 4 ? 2 : P == 2 ? 1 : 0; }

static int build_program(std::vector<RefProgram> refs, saf_program **out) {
    std::unique_ptr<saf_program> p(new (std::nothrow) saf_program());
    if (!p) return fail(SAF_ERR_ALLOC, "out of host memory");
    p->refs = std::move(refs);
    std::vector<const RefProgram *> ptrs;
    for (auto &r : p->refs) ptrs.push_back(&r);
    std::string err;
    int rc = lower(ptrs, p->dev, err);
    if (rc != SAF_OK) return fail(rc, err);
    if (p->dev.param_count > (uint32_t)MAX_COLS || p->dev.result_field.size() > (size_t)MAX_OUTS) {
        return fail(SAF_ERR_UNSUPPORTED, "programs with more than 32 parameters or 32 results are not supported");
    }
    // pack the widest layout now so that encoding problems surface at creation time
    rc = pack_program(p->dev, 4, p->packed[2], err);
    if (rc != SAF_OK) return fail(rc, err);
    p->packed_ready[2] = true;
    *out = p.release();
    return SAF_OK;
}

extern "C" int saf_program_create(const uint32_t *flat, size_t n_words, const double *consts, size_t n_consts,
                                  const uint32_t *pool, size_t n_pool, uint32_t param_count,
                                  uint32_t workspace_size, const uint32_t *result_regs, uint32_t n_results,
                                  saf_program **out) {
    if (!out) return fail(SAF_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!flat || n_words == 0) return fail(SAF_ERR_INVALID_ARGUMENT, "empty bytecode");
    if ((n_consts && !consts) || (n_pool && !pool) || !result_regs || n_results == 0)
        return fail(SAF_ERR_INVALID_ARGUMENT, "NULL program component");
    RefProgram r;
    r.flat.assign(flat, flat + n_words);
    if (n_consts) r.consts.assign(consts, consts + n_consts);
    if (n_pool) r.pool.assign(pool, pool + n_pool);
    r.param_count = param_count;
    r.workspace_size = workspace_size;
    r.result_regs.assign(result_regs, result_regs + n_results);
    std::vector<RefProgram> v;
    v.push_back(std::move(r));
    return build_program(std::move(v), out);
}

extern "C" int saf_program_fuse(const saf_program *const *programs, size_t n_programs, saf_program **out) {
    if (!out) return fail(SAF_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!programs || n_programs == 0) return fail(SAF_ERR_INVALID_ARGUMENT, "no programs to fuse");
    std::vector<RefProgram> v;
    for (size_t i = 0; i < n_programs; ++i) {
        if (!programs[i]) return fail(SAF_ERR_INVALID_ARGUMENT, "NULL program in fuse list");
        for (const auto &r : programs[i]->refs) v.push_back(r);
    }
    return build_program(std::move(v), out);
}

namespace saf { namespace cc {
int compile(const std::vector<std::string> &exprs, const std::vector<std::string> &params, RefProgram &out, std::string &err);
} } // namespace saf::cc

extern "C" int saf_compile_exprs(const char *const *exprs, size_t n_exprs, const char *const *params, size_t n_params,
                                 saf_program **out) {
    if (!out) return fail(SAF_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!exprs || n_exprs == 0) return fail(SAF_ERR_INVALID_ARGUMENT, "no expression to compile");
    if (n_params && !params) return fail(SAF_ERR_INVALID_ARGUMENT, "params is NULL");
    std::vector<std::string> es, ps;
    for (size_t i = 0; i < n_exprs; ++i) {
        if (!exprs[i]) return fail(SAF_ERR_INVALID_ARGUMENT, "NULL expression");
        es.emplace_back(exprs[i]);
    }
    for (size_t i = 0; i < n_params; ++i) {
        if (!params[i]) return fail(SAF_ERR_INVALID_ARGUMENT, "NULL parameter name");
        ps.emplace_back(params[i]);
    }
    RefProgram r;
    std::string err;
    const int rc = saf::cc::compile(es, ps, r, err);
    if (rc != SAF_OK) return fail(rc, err);
    std::vector<RefProgram> v;
    v.push_back(std::move(r));
    return build_program(std::move(v), out);
}

extern "C" int saf_program_reference_payload(const saf_program *p, size_t index, uint32_t *flat, size_t cap_words,
                                             size_t *n_words, double *consts, size_t cap_consts, size_t *n_consts,
                                             uint32_t *arg_pool, size_t cap_pool, size_t *n_pool, uint32_t *param_count,
                                             uint32_t *workspace_size, uint32_t *result_regs, size_t cap_results,
                                             size_t *n_results) {
    if (!p) return fail(SAF_ERR_INVALID_ARGUMENT, "program is NULL");
    if (index >= p->refs.size()) return fail(SAF_ERR_INVALID_ARGUMENT, "component index out of range");
    const RefProgram &r = p->refs[index];
    if (n_words) *n_words = r.flat.size();
    if (n_consts) *n_consts = r.consts.size();
    if (n_pool) *n_pool = r.pool.size();
    if (n_results) *n_results = r.result_regs.size();
    if (param_count) *param_count = r.param_count;
    if (workspace_size) *workspace_size = r.workspace_size;
    if (flat) std::memcpy(flat, r.flat.data(), std::min(cap_words, r.flat.size()) * 4);
    if (consts) std::memcpy(consts, r.consts.data(), std::min(cap_consts, r.consts.size()) * 8);
    if (arg_pool) std::memcpy(arg_pool, r.pool.data(), std::min(cap_pool, r.pool.size()) * 4);
    if (result_regs) std::memcpy(result_regs, r.result_regs.data(), std::min(cap_results, r.result_regs.size()) * 4);
    return SAF_OK;
}

extern "C" void saf_program_destroy(saf_program *p) {
    if (!p) return;
    for (auto &sp : p->spec)
        for (auto &kv : sp.modules) {
            int prev = 0;
            if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(kv.first) == cudaSuccess) {
                spec_module_unload(&kv.second);
                cudaSetDevice(prev);
            }
        }
    for (auto &kv : p->copies) {
        int prev = 0;
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(kv.first.first) == cudaSuccess) {
            cudaFree(kv.second);
            cudaSetDevice(prev);
        }
    }
    delete p;
}

// ---- content-addressed program cache (saf_program_intern) ------------------------------------------------------
namespace {
std::mutex g_intern_mu;
std::multimap<uint64_t, saf_program *> g_interned;
constexpr size_t kInternCapacity = 4096;

uint64_t fnv1a(uint64_t h, const void *data, size_t n) {
    const unsigned char *b = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < n; ++i) {
        h ^= b[i];
        h *= 0x100000001b3ull;
    }
    return h;
}
uint64_t payload_hash(const RefProgram &r) {
    uint64_t h = 0xcbf29ce484222325ull;
    const uint64_t sizes[4] = {r.flat.size(), r.consts.size(), r.pool.size(), r.result_regs.size()};
    h = fnv1a(h, sizes, sizeof sizes);
    h = fnv1a(h, r.flat.data(), r.flat.size() * 4);
    h = fnv1a(h, r.consts.data(), r.consts.size() * 8);
    h = fnv1a(h, r.pool.data(), r.pool.size() * 4);
    h = fnv1a(h, r.result_regs.data(), r.result_regs.size() * 4);
    h = fnv1a(h, &r.param_count, 4);
    return fnv1a(h, &r.workspace_size, 4);
}
bool same_payload(const RefProgram &a, const RefProgram &b) {
    return a.param_count == b.param_count && a.workspace_size == b.workspace_size && a.flat == b.flat && a.pool == b.pool &&
           a.result_regs == b.result_regs && a.consts.size() == b.consts.size() &&
           (a.consts.empty() || std::memcmp(a.consts.data(), b.consts.data(), a.consts.size() * 8) == 0); // bitwise: NaN, -0.0
}
} // namespace

extern "C" int saf_program_intern(const uint32_t *flat, size_t n_words, const double *consts, size_t n_consts,
                                  const uint32_t *pool, size_t n_pool, uint32_t param_count,
                                  uint32_t workspace_size, const uint32_t *result_regs, uint32_t n_results,
                                  saf_program **out) {
    if (!out) return fail(SAF_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!flat || n_words == 0) return fail(SAF_ERR_INVALID_ARGUMENT, "empty bytecode");
    if ((n_consts && !consts) || (n_pool && !pool) || !result_regs || n_results == 0)
        return fail(SAF_ERR_INVALID_ARGUMENT, "NULL program component");
    RefProgram r;
    r.flat.assign(flat, flat + n_words);
    if (n_consts) r.consts.assign(consts, consts + n_consts);
    if (n_pool) r.pool.assign(pool, pool + n_pool);
    r.param_count = param_count;
    r.workspace_size = workspace_size;
    r.result_regs.assign(result_regs, result_regs + n_results);
    const uint64_t h = payload_hash(r);
    std::lock_guard<std::mutex> lk(g_intern_mu); // lowering under the lock: concurrent callers of one program wait for one lowering
    for (auto range = g_interned.equal_range(h); range.first != range.second; ++range.first)
        if (range.first->second->refs.size() == 1 && same_payload(range.first->second->refs[0], r)) {
            *out = range.first->second;
            return SAF_OK;
        }
    if (g_interned.size() >= kInternCapacity)
        return fail(SAF_ERR_ALLOC, "program cache is full (4096 programs): call saf_program_cache_clear");
    std::vector<RefProgram> v;
    v.push_back(std::move(r));
    saf_program *p = nullptr;
    const int rc = build_program(std::move(v), &p);
    if (rc != SAF_OK) return rc;
    g_interned.emplace(h, p);
    *out = p;
    return SAF_OK;
}

extern "C" void saf_program_cache_clear(void) {
    std::lock_guard<std::mutex> lk(g_intern_mu);
    for (auto &kv : g_interned) saf_program_destroy(kv.second);
    g_interned.clear();
}

extern "C" int saf_program_get_info(const saf_program *p, saf_program_info *info) {
    if (!p || !info) return fail(SAF_ERR_INVALID_ARGUMENT, "NULL argument");
    info->param_count = p->dev.param_count;
    info->n_results = (uint32_t)p->dev.result_field.size();
    info->n_ref_instructions = p->dev.n_ref_instructions;
    info->n_dev_instructions = p->dev.n_dev_instructions;
    info->n_slots = p->dev.n_slots;
    info->n_consts = (uint32_t)p->dev.consts.size();
    info->n_acc_forwards = p->dev.n_acc_forwards;
    info->quirk_flags = p->dev.quirk_flags;
    info->fp64_slots = p->dev.fp64_slots;
    uint32_t used = 0;
    for (uint16_t s : p->dev.param_slot) used += (s != NO_SLOT);
    info->bytes_per_point = 8.0 * (double)(used + p->dev.result_field.size());
    return SAF_OK;
}

extern "C" size_t saf_program_disassemble(const saf_program *p, char *buf, size_t cap) {
    if (!p) return 0;
    std::string s = disassemble(p->dev);
    if (buf && cap) {
        size_t n = std::min(cap - 1, s.size());
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

extern "C" int saf_program_export(const saf_program *p, uint64_t *code, size_t cap_code, size_t *n_code,
                                  double *consts, size_t cap_consts, size_t *n_consts, uint16_t *param_slot,
                                  uint32_t *result_field) {
    if (!p) return fail(SAF_ERR_INVALID_ARGUMENT, "program is NULL");
    if (n_code) *n_code = p->dev.code.size();
    if (n_consts) *n_consts = p->dev.consts.size();
    if (code) std::memcpy(code, p->dev.code.data(), std::min(cap_code, p->dev.code.size()) * 8);
    if (consts) std::memcpy(consts, p->dev.consts.data(), std::min(cap_consts, p->dev.consts.size()) * 8);
    if (param_slot) std::copy(p->dev.param_slot.begin(), p->dev.param_slot.end(), param_slot);
    if (result_field) std::copy(p->dev.result_field.begin(), p->dev.result_field.end(), result_field);
    return SAF_OK;
}


// ---- specialised kernels (spec.h) ------------------------------------------------------------------------------------
// SAF_SPECIALIZE: 0 = never launch a specialised kernel (the interpreter runs even after saf_program_specialize);
// 1 = specialise every program at its first launch (how the whole GPU test-suite is run against this machine);
// auto = specialise a program once the work it has done on the interpreter says it is here to stay (below);
// unset = specialised kernels run for the programs saf_program_specialize was called on.
static int specialize_mode() {
    static const int mode = []() {
        const char *v = getenv("SAF_SPECIALIZE");
        return !v || !*v ? -1 : v[0] == '0' ? 0 : v[0] == 'a' ? 2 : 1;
    }();
    return mode;
}
// auto mode: instruction-points (device instructions x points x iterations) a handle must have run on the interpreter
// before it is specialised at its next launch: 2e10 is 50-100 ms of interpreter time, the order of one compilation
constexpr double kAutoSpecializeWork = 2e10;

static int specialize_locked(saf_program *p, int variant = 0) { // p->mu held
    saf_program::Spec &sp = p->spec[variant];
    if (sp.attempted) return sp.rc;
    sp.attempted = true;
    const auto t0 = std::chrono::steady_clock::now();
    sp.shape = spec_shape_for(p->dev, variant == 1);
    sp.rc = spec_generate(p->dev, sp.shape, sp.text, sp.can_iterate, sp.has_prefetch_entry, sp.err);
    const auto t1 = std::chrono::steady_clock::now();
    sp.generate_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (sp.rc != SAF_OK) return sp.rc;
    sp.rc = jit_compile(sp.text, sp.cubin, sp.err);
    sp.compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    return sp.rc;
}

extern "C" int saf_program_specialize(saf_program *p) {
    if (!p) return fail(SAF_ERR_INVALID_ARGUMENT, "NULL program");
    std::lock_guard<std::mutex> lk(p->mu);
    const int rc = specialize_locked(p);
    if (rc != SAF_OK) return fail(rc, "specialisation failed: " + p->spec[0].err);
    return SAF_OK;
}

extern "C" int saf_program_specialized_info(const saf_program *p, saf_specialized_info *info) {
    if (!p || !info) return fail(SAF_ERR_INVALID_ARGUMENT, "NULL argument");
    saf_program *mp = const_cast<saf_program *>(p);
    std::lock_guard<std::mutex> lk(mp->mu);
    const saf_program::Spec &sp = p->spec[0];
    std::memset(info, 0, sizeof *info);
    info->compiled = sp.attempted && sp.rc == SAF_OK;
    info->can_iterate = sp.can_iterate;
    info->points_per_thread = sp.shape.points_per_thread;
    info->block_threads = sp.shape.block;
    info->min_blocks = sp.shape.min_blocks;
    info->generate_ms = sp.generate_ms;
    info->compile_ms = sp.compile_ms;
    info->cubin_bytes = sp.cubin.size();
    info->source_bytes = sp.text.size();
    info->registers = -1;
    info->local_bytes = -1;
    info->resident_blocks = -1;
    info->small_registers = -1;
    info->prefetch_registers = -1;
    info->has_prefetch_entry = sp.has_prefetch_entry;
    const saf_program::Spec &sm = p->spec[1];
    info->small_compiled = sm.attempted && sm.rc == SAF_OK;
    info->small_points_per_thread = sm.shape.points_per_thread;
    info->small_block_threads = sm.shape.block;
    info->small_compile_ms = sm.compile_ms;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        auto it = sp.modules.find(dev);
        if (it != sp.modules.end()) {
            info->registers = it->second.registers;
            info->local_bytes = it->second.local_bytes;
            info->resident_blocks = it->second.resident_blocks;
            info->prefetch_registers = it->second.function_pf ? it->second.registers_pf : -1;
        }
        it = sm.modules.find(dev);
        if (it != sm.modules.end()) info->small_registers = it->second.registers;
    }
    return SAF_OK;
}

extern "C" size_t saf_program_specialized_source(const saf_program *p, int small_batch, char *buf, size_t cap) {
    if (!p) return 0;
    saf_program *mp = const_cast<saf_program *>(p);
    std::lock_guard<std::mutex> lk(mp->mu);
    const std::string &s = p->spec[small_batch ? 1 : 0].text;
    if (buf && cap) {
        const size_t n = std::min(cap - 1, s.size());
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

extern "C" int saf_specialize_available(void) {
    std::string why;
    if (jit_available(why)) return 1;
    g_last_error = why;
    return 0;
}

// The specialised kernel of `prog` for a batch of n_points on the current device, or nullptr when the interpreter
// should run.
static int spec_for_launch(const saf_program *prog, uint64_t iters, size_t n_points, const SpecModule **out, SpecShape *shape) {
    *out = nullptr;
    const int mode = specialize_mode();
    if (mode == 0) return SAF_OK;
    saf_program *mp = const_cast<saf_program *>(prog);
    std::lock_guard<std::mutex> lk(mp->mu);
    if (!mp->spec[0].attempted) {
        if (mode == 2) {
            if (mp->interpreted_work < kAutoSpecializeWork) { // this launch still goes to the interpreter, and counts
                mp->interpreted_work += (double)mp->dev.n_dev_instructions * (double)n_points * (double)(iters ? iters : 1);
                return SAF_OK;
            }
        } else if (mode != 1) {
            return SAF_OK;
        }
        // programs that cannot be specialised (too long, no NVRTC) keep running on the interpreter
        if (specialize_locked(mp) != SAF_OK) return SAF_OK;
    }
    if (mp->spec[0].rc != SAF_OK) return SAF_OK;
    int dev = 0, sms = 0;
    SAF_CUDA(cudaGetDevice(&dev));
    SAF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int variant = spec_batch_is_small(mp->spec[0].shape, n_points, sms) ? 1 : 0;
    if (const char *v = getenv("SAF_SPEC_VARIANT")) variant = v[0] == 's' ? 1 : v[0] == 'l' ? 0 : variant; // test knob
    // the small-batch shape is compiled the first time a small batch arrives; if that fails the large shape runs
    if (variant == 1 && specialize_locked(mp, 1) != SAF_OK) variant = 0;
    saf_program::Spec &sp = mp->spec[variant];
    if (iters > 1 && !sp.can_iterate) return SAF_OK;
    auto it = sp.modules.find(dev);
    if (it == sp.modules.end()) {
        SpecModule m;
        cudaError_t e = spec_module_load(sp.cubin.data(), sp.has_prefetch_entry, &m);
        if (e != cudaSuccess) return cuda_fail(e, "loading the specialised kernel");
        if (m.max_threads < sp.shape.block) {
            spec_module_unload(&m);
            return fail(SAF_ERR_CUDA, "specialised kernel cannot run " + std::to_string(sp.shape.block) + " threads per block");
        }
        e = spec_module_occupancy(&m, sp.shape.block);
        if (e != cudaSuccess) {
            spec_module_unload(&m);
            return cuda_fail(e, "occupancy of the specialised kernel");
        }
        it = sp.modules.emplace(dev, m).first;
    }
    *out = &it->second;
    *shape = sp.shape;
    return SAF_OK;
}

constexpr uint64_t kMaxItersPerLaunch = 1ull << 30;

// ---- launch -------------------------------------------------------------------------------------------
static int packed_for(saf_program *p, int P, const PackedProgram **out) {
    const int li = layout_index(P);
    std::lock_guard<std::mutex> lk(p->mu);
    if (!p->packed_ready[li]) {
        std::string err;
        int rc = P == LAYOUT_R ? pack_program_r(p->dev, p->packed[li], err) : pack_program(p->dev, P, p->packed[li], err);
        if (P == LAYOUT_R) p->r_pack_rc = rc;
        if (rc != SAF_OK && P != LAYOUT_R) return fail(rc, err);
        if (rc != SAF_OK) g_last_error = err;
        p->packed_ready[li] = true;
    }
    if (P == LAYOUT_R && p->r_pack_rc != SAF_OK) {
        return p->r_pack_rc;
    }
    *out = &p->packed[li];
    return SAF_OK;
}

static int ensure_device_image(saf_program *p, int device, const PackedProgram &pk, const unsigned char **out) {
    std::lock_guard<std::mutex> lk(p->mu);
    const auto key = std::make_pair(device, pk.layout_id);
    auto it = p->copies.find(key);
    if (it != p->copies.end()) {
        *out = it->second;
        return SAF_OK;
    }
    unsigned char *dptr = nullptr;
    SAF_CUDA(cudaMalloc(&dptr, pk.image.size() * 8));
    // The image is pageable host memory: cudaMemcpy may return once the bytes are staged, before the DMA has reached
    // device memory, and the first launch goes onto a non-blocking stream that is not ordered after the legacy
    // default stream.  Wait for the copy itself before the pointer is published.
    cudaError_t ce = cudaMemcpy(dptr, pk.image.data(), pk.image.size() * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(cudaStreamLegacy);
    if (ce != cudaSuccess) {
        cudaFree(dptr);
        return cuda_fail(ce, "program image upload");
    }
    p->copies[key] = dptr;
    *out = dptr;
    return SAF_OK;
}

// cols/outs are DEVICE pointers here; lens are the device-side lengths (already chunk-relative).
static int launch_on_device(const saf_program *prog, const double *const *cols, const size_t *lens, size_t n_cols,
                            double *const *outs, size_t n_points, uint64_t iters, unsigned flags,
                            cudaStream_t stream) {
    if (n_points == 0) return SAF_OK;
    const DeviceProgram &dp = prog->dev;
    {
        const SpecModule *sm = nullptr;
        SpecShape ss;
        const int src = spec_for_launch(prog, iters, n_points, &sm, &ss);
        if (src != SAF_OK) return src;
        if (sm != nullptr) {
            SpecDesc sd;
            std::memset(&sd, 0, sizeof sd);
            sd.n_points = n_points;
            sd.iters = iters ? iters : 1;
            for (uint32_t j = 0; j < dp.param_count; ++j) {
                if (j < n_cols && cols[j] != nullptr) {
                    sd.cols[j] = cols[j];
                    sd.col_len[j] = lens[j];
                }
            }
            for (size_t k = 0; k < dp.result_field.size(); ++k) sd.outs[k] = outs[k];
            const unsigned long long tile = (unsigned long long)ss.block * (unsigned long long)ss.points_per_thread;
            const unsigned long long n_tiles = (n_points + tile - 1) / tile;
            const bool pf = sd.iters == 1 && sm->function_pf != nullptr; // single pass: the entry point with next-tile prefetch
            unsigned long long grid = (unsigned long long)sm->sms * (unsigned long long)(pf ? sm->resident_blocks_pf : sm->resident_blocks);
            if (grid > n_tiles) grid = n_tiles;
            cudaError_t e = spec_launch(*sm, pf, sd, (int)grid, ss.block, stream);
            if (e != cudaSuccess) return cuda_fail(e, "specialised kernel launch");
            g_launches.fetch_add(1, std::memory_order_relaxed);
            return SAF_OK;
        }
    }
    LaunchDesc d;
    std::memset(&d, 0, sizeof d);
    d.n_points = n_points;
    d.iters = iters ? iters : 1;
    d.n_params = dp.param_count;
    d.n_results = (uint32_t)dp.result_field.size();
    d.flags = flags;
    bool aligned = true;
    for (uint32_t j = 0; j < dp.param_count; ++j) {
        if (j < n_cols && cols[j] != nullptr) {
            d.cols[j] = cols[j];
            d.col_len[j] = lens[j];
            if (dp.param_slot[j] != NO_SLOT && ((uintptr_t)cols[j] & 15u)) aligned = false;
        } else {
            d.cols[j] = nullptr; // missing parameter column -> 0.0 (scalar.rs:192-197)
            d.col_len[j] = 0;
        }
    }
    for (uint32_t k = 0; k < d.n_results; ++k) {
        d.outs[k] = outs[k];
        if ((uintptr_t)outs[k] & 15u) aligned = false;
    }
    int device = 0;
    SAF_CUDA(cudaGetDevice(&device));
    static const int forced_ppt = []() {
        const char *v = getenv("SAF_POINTS_PER_THREAD"); // tuning knob, see DESIGN.md
        return v ? atoi(v) : 0;
    }();
    int P = 1;
    cudaError_t e = choose_points_per_thread(device, dp.n_slots, n_points, aligned, forced_ppt, &P);
    if (e == cudaErrorInvalidConfiguration)
        return fail(SAF_ERR_UNSUPPORTED, "program keeps more live values than fit the on-chip register file");
    if (e != cudaSuccess) return cuda_fail(e, "choose_points_per_thread");
    const PackedProgram *pk = nullptr;
    int rc = SAF_OK;
    // Register machine (ruop.h): live values in real registers, four points per thread.  Opt-in (SAF_MACHINE=r):
    // measured slower than the shared-memory machine on all five named workloads this round -- two-address code
    // needs 1.5-1.8x the micro-ops, and a dispatch (indexed constant load + three taken branches) costs more than
    // the shared-memory operand loads it saves (profiles/r01_h_*).  SAF_MACHINE=s forces the default.
    static const int forced_machine = []() {
        const char *v = getenv("SAF_MACHINE");
        return !v ? 0 : v[0] == 'r' ? 1 : v[0] == 's' ? 2 : 0;
    }();
    bool use_r = false;
    if (aligned && forced_machine == 1 && dp.consts.size() * 8 + 64 <= K_SMEM_MAX) {
        const PackedProgram *rp = nullptr;
        LaunchShape rs;
        if (packed_for(const_cast<saf_program *>(prog), LAYOUT_R, &rp) == SAF_OK && (d.iters == 1 || rp->can_iterate) &&
            choose_shape(device, rp->n_slots, rp->code_off, rp->code_bytes, n_points, R_POINTS, true, &rs) == cudaSuccess &&
            rs.k_in_smem) {
            pk = rp;
            use_r = true;
            P = R_POINTS;
        }
    }
    if (!use_r) rc = packed_for(const_cast<saf_program *>(prog), P, &pk);
    if (rc != SAF_OK) return rc;
    // Programs with many live values fit two 128-thread blocks per SM at most (shared-memory register file): one
    // wide block holds more warps in the same shared memory (SAF_WIDE=0 disables: tuning knob).
    int block_threads = BLOCK_THREADS;
    static const bool wide_ok = []() {
        const char *v = getenv("SAF_WIDE");
        return !(v && v[0] == '0');
    }();
    if (!use_r && (P == 2 || P == 4) && wide_ok) {
        const int t = wide_block_threads(device, pk->n_slots, pk->code_off, pk->code_bytes, n_points, P);
        if (t > 0) {
            saf_program *mp = const_cast<saf_program *>(prog);
            std::lock_guard<std::mutex> lk(mp->mu);
            const int key = 1000 * P + t;
            auto it = mp->packed_wide.find(key);
            if (it == mp->packed_wide.end()) {
                std::string err;
                PackedProgram w;
                if (pack_program(mp->dev, P, w, err, t) == SAF_OK) it = mp->packed_wide.emplace(key, std::move(w)).first;
            }
            if (it != mp->packed_wide.end()) {
                pk = &it->second; // map entries never move: valid after the lock is released
                block_threads = t;
            }
        }
    }
    d.code_off = pk->code_off;
    d.n_slots = pk->n_slots;
    d.acc_off = pk->acc_off;
    d.x0_off = pk->x0_off;
    d.x1_off = pk->x1_off;
    for (uint32_t j = 0; j < dp.param_count; ++j) d.param_off[j] = pk->param_off[j];
    const bool even = pk->double_buffered && (d.iters % 2ull == 0ull);
    for (uint32_t k = 0; k < d.n_results; ++k) {
        d.result_off[k] = even ? pk->result_off_even[k] : pk->result_off[k];
        if (pk->result_is_k[k]) d.result_k_mask |= 1u << k;
    }
    d.n_feedback = (pk->double_buffered || use_r) ? 0u : dp.param_count; // parameters hot_run copies back between iterations
    rc = ensure_device_image(const_cast<saf_program *>(prog), device, *pk, &d.image_global);
    if (rc != SAF_OK) return rc;
    d.code_bytes = pk->code_bytes;
    // streaming (kernel.cu): a single pass whose columns are all present, full length and 16-byte aligned, whose results
    // all live in slots, and whose code fits one window
    bool stream_ok = aligned && d.iters == 1 && !use_r && P >= 2 && d.result_k_mask == 0;
    for (uint32_t j = 0; j < dp.param_count && stream_ok; ++j)
        if (dp.param_slot[j] != NO_SLOT && (d.cols[j] == nullptr || d.col_len[j] < n_points)) stream_ok = false;
    LaunchShape shape;
    e = choose_shape(device, pk->n_slots, pk->code_off, pk->code_bytes, n_points, P, use_r, &shape, block_threads, stream_ok);
    if (e == cudaErrorInvalidConfiguration)
        return fail(SAF_ERR_UNSUPPORTED, "program keeps more live values than fit the on-chip register file");
    if (e != cudaSuccess) return cuda_fail(e, "choose_shape");
    d.n_rf = (uint32_t)shape.n_rf;
    e = launch_interpreter(d, shape, stream);
    if (e != cudaSuccess) return cuda_fail(e, "interpreter launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SAF_OK;
}

extern "C" int saf_program_export_packed(const saf_program *p, int points_per_thread, uint64_t *image, size_t cap_words,
                                         size_t *n_words, uint32_t *code_off, uint32_t *n_slots, int32_t *x_off,
                                         int32_t *param_off, int32_t *result_off, uint8_t *result_is_k,
                                         int32_t *result_off_even, int *double_buffered) {
    if (!p) return fail(SAF_ERR_INVALID_ARGUMENT, "program is NULL");
    const PackedProgram *pk = nullptr;
    PackedProgram wide; // 1000 * P + threads: the wide one-block-per-SM layout (P = 2 or 4), packed on the spot
    int rc;
    const int wide_p = points_per_thread / 1000, wide_t = points_per_thread % 1000;
    if ((wide_p == 2 || wide_p == 4) && wide_t > 0 && wide_t <= WIDE_MAX_THREADS && wide_t % 32 == 0) {
        std::string err;
        rc = pack_program(p->dev, wide_p, wide, err, wide_t);
        if (rc != SAF_OK) return fail(rc, err);
        pk = &wide;
    } else {
        if (points_per_thread != 1 && points_per_thread != 2 && points_per_thread != 4 && points_per_thread != 8 &&
            points_per_thread != LAYOUT_R)
            return fail(SAF_ERR_INVALID_ARGUMENT,
                        "points_per_thread must be 1, 2, 4, 8, 104 (register machine) or 1000 * P + threads (wide blocks, P = 2 or 4)");
        rc = packed_for(const_cast<saf_program *>(p), points_per_thread, &pk);
        if (rc != SAF_OK) return rc;
    }
    if (n_words) *n_words = pk->image.size();
    if (image) std::memcpy(image, pk->image.data(), std::min(cap_words, pk->image.size()) * 8);
    if (code_off) *code_off = pk->code_off;
    if (n_slots) *n_slots = pk->n_slots;
    if (x_off) {
        x_off[0] = pk->x0_off;
        x_off[1] = pk->x1_off;
        x_off[2] = pk->acc_off;
    }
    if (param_off) std::copy(pk->param_off.begin(), pk->param_off.end(), param_off);
    if (result_off) std::copy(pk->result_off.begin(), pk->result_off.end(), result_off);
    if (result_off_even) std::copy(pk->result_off_even.begin(), pk->result_off_even.end(), result_off_even);
    if (double_buffered) *double_buffered = pk->double_buffered ? 1 : 0;
    if (result_is_k) std::copy(pk->result_is_k.begin(), pk->result_is_k.end(), result_is_k);
    return SAF_OK;
}

// run_chunked_evaluator's argument checks (batch.rs:44-50) / eval_batch's (scalar.rs:157-159)
static int check_columns(const saf_program *prog, const double *const *cols, const size_t *lens, size_t n_cols,
                         double *const *outs, size_t n_points, unsigned flags) {
    if (!prog) return fail(SAF_ERR_INVALID_ARGUMENT, "program is NULL");
    if (n_cols && (!cols || !lens)) return fail(SAF_ERR_INVALID_ARGUMENT, "cols/col_lens is NULL");
    if (n_points && !outs) return fail(SAF_ERR_INVALID_ARGUMENT, "outs is NULL");
    if (!(flags & SAF_EVAL_BROADCAST)) {
        for (size_t j = 0; j < n_cols; ++j)
            if (lens[j] != n_points)
                return fail(SAF_ERR_COLUMN_LENGTH_MISMATCH, "All evaluation columns must have the same length");
    } else {
        for (size_t j = 0; j < n_cols && j < prog->dev.param_count; ++j)
            if (n_points && lens[j] == 0)
                return fail(SAF_ERR_INVALID_ARGUMENT, "broadcast column is empty");
    }
    for (size_t j = 0; j < n_cols && j < prog->dev.param_count; ++j)
        if (n_points && lens[j] && !cols[j]) return fail(SAF_ERR_INVALID_ARGUMENT, "column pointer is NULL");
    for (size_t k = 0; n_points && k < prog->dev.result_field.size(); ++k)
        if (!outs[k]) return fail(SAF_ERR_INVALID_ARGUMENT, "output pointer is NULL");
    return SAF_OK;
}

extern "C" int saf_eval_device(const saf_program *prog, const double *const *cols, const size_t *lens, size_t n_cols,
                               double *const *outs, size_t n_points, unsigned flags, void *stream) {
    int rc = check_columns(prog, cols, lens, n_cols, outs, n_points, flags);
    if (rc != SAF_OK) return rc;
    return launch_on_device(prog, cols, lens, n_cols, outs, n_points, 1, flags, (cudaStream_t)stream);
}

extern "C" int saf_iterate_device(const saf_program *prog, double *const *state, size_t n_state, size_t n_points,
                                  uint64_t iters, void *stream) {
    if (!prog || (n_points && !state)) return fail(SAF_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_state != prog->dev.param_count || n_state != prog->dev.result_field.size())
        return fail(SAF_ERR_INVALID_ARGUMENT, "iterate needs n_state == param_count == n_results");
    if (iters == 0 || n_points == 0) return SAF_OK;
    std::vector<size_t> lens(n_state, n_points);
    std::vector<const double *> cols(state, state + n_state);
    // the kernel counts iterations in 32 bits: longer runs are a sequence of launches (the state is in the columns)
    while (iters > 0) {
        const uint64_t now = std::min<uint64_t>(iters, kMaxItersPerLaunch);
        int rc = launch_on_device(prog, cols.data(), lens.data(), n_state, state, n_points, now, 0, (cudaStream_t)stream);
        if (rc != SAF_OK) return rc;
        iters -= now;
    }
    return SAF_OK;
}

// ---- host-buffer path -----------------------------------------------------------------------------------
namespace {

// Small per-process cache of device scratch so that repeated host calls (the reference's callers
// invoke eval_f64 once per simulation step) do not pay cudaMalloc each time.
struct ScratchPool {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks[64];
    void *acquire(int dev, size_t bytes, cudaError_t *err) {
        {
            std::lock_guard<std::mutex> lk(mu);
            auto &m = free_blocks[dev & 63];
            auto it = m.lower_bound(bytes);
            if (it != m.end() && it->first <= bytes * 2 + (1 << 20)) {
                void *p = it->second;
                sizes[p] = it->first;
                m.erase(it);
                return p;
            }
        }
        void *p = nullptr;
        *err = cudaMalloc(&p, bytes);
        if (*err != cudaSuccess) return nullptr;
        std::lock_guard<std::mutex> lk(mu);
        sizes[p] = bytes;
        return p;
    }
    void release(int dev, void *p) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu);
        free_blocks[dev & 63].emplace(sizes[p], p);
    }
    std::map<void *, size_t> sizes;
};
ScratchPool g_pool;

struct StreamSet {
    cudaStream_t s[3] = {nullptr, nullptr, nullptr};
};
std::mutex g_stream_mu;
std::map<int, std::vector<StreamSet>> g_stream_cache;

int acquire_streams(int dev, StreamSet *out) {
    {
        std::lock_guard<std::mutex> lk(g_stream_mu);
        auto &v = g_stream_cache[dev];
        if (!v.empty()) {
            *out = v.back();
            v.pop_back();
            return SAF_OK;
        }
    }
    for (auto &s : out->s) SAF_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return SAF_OK;
}
void release_streams(int dev, const StreamSet &s) {
    std::lock_guard<std::mutex> lk(g_stream_mu);
    g_stream_cache[dev].push_back(s);
}

constexpr size_t kChunkPoints = 1u << 21; // 2 Mi points: 16 MiB per column per chunk

// ---- pageable host buffers --------------------------------------------------------------------------------------
// Real callers hold pageable memory (a Rust Vec<f64>, a NumPy array: src/bindings/python/parallel.rs:198-223 passes
// the array's own buffer).  cudaMemcpyAsync on pageable memory is staged by the driver and SYNCHRONOUS with respect to
// the host, which serialises the three-stream pipeline below.  Pageable columns are therefore staged here, chunk by
// chunk, through a pooled ring of pinned buffers: the host thread (helped by a few copy threads) moves chunk k + 1 into
// pinned memory while the GPU copies and evaluates chunk k.
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;
    std::map<void *, size_t> sizes;
    size_t cached = 0;
    static constexpr size_t kMaxCached = (size_t)2 << 30;
    void *acquire(size_t bytes, cudaError_t *err) {
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_blocks.lower_bound(bytes);
            if (it != free_blocks.end() && it->first <= bytes * 2 + (1 << 20)) {
                void *p = it->second;
                cached -= it->first;
                free_blocks.erase(it);
                return p;
            }
        }
        void *p = nullptr;
        *err = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
        if (*err != cudaSuccess) return nullptr;
        std::lock_guard<std::mutex> lk(mu);
        sizes[p] = bytes;
        return p;
    }
    void release(void *p) {
        if (!p) return;
        std::unique_lock<std::mutex> lk(mu);
        const size_t b = sizes[p];
        if (cached + b > kMaxCached) {
            sizes.erase(p);
            lk.unlock();
            cudaFreeHost(p);
            return;
        }
        cached += b;
        free_blocks.emplace(b, p);
    }
};
PinnedPool g_pinned;

bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError(); // older runtimes report unregistered host memory as an error
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

struct CopyJob { void *dst; const void *src; size_t bytes; };
// Persistent copy workers: a chunk is staged by two parallel_copy calls (in, out), a 100 M-point batch by a hundred;
// spawning and joining five threads per call cost 0.1-0.3 ms each time -- as much as the copy of a 16 MB chunk itself.
class CopyPool {
  public:
    void run(std::vector<std::function<void()>> &tasks) { // runs all tasks, the last one on the calling thread
        if (tasks.empty()) return;
        struct Done {
            std::mutex mu;
            std::condition_variable cv;
            size_t left;
        } done;
        done.left = tasks.size() - 1;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (pid_ != getpid()) { // forked child: the parent's workers do not exist here
                pid_ = getpid();
                n_workers_ = 0;
                queue_.clear();
            }
            while (n_workers_ + 1 < tasks.size() && n_workers_ < kMaxWorkers) {
                std::thread([this]() { work(); }).detach();
                ++n_workers_;
            }
            for (size_t i = 0; i + 1 < tasks.size(); ++i) {
                std::function<void()> *t = &tasks[i];
                queue_.push_back([t, &done]() {
                    (*t)();
                    std::lock_guard<std::mutex> lk(done.mu);
                    if (--done.left == 0) done.cv.notify_one();
                });
            }
        }
        cv_.notify_all();
        tasks.back()();
        std::unique_lock<std::mutex> lk(done.mu);
        done.cv.wait(lk, [&]() { return done.left == 0; });
    }

  private:
    static constexpr size_t kMaxWorkers = 16;
    void work() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&]() { return !queue_.empty(); });
                f = std::move(queue_.front());
                queue_.pop_front();
            }
            f();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> queue_;
    size_t n_workers_ = 0;
    pid_t pid_ = getpid();
};
// never destroyed: its detached workers wait on its condition variable until the process ends
CopyPool &g_copy_pool = *new CopyPool();

// host-to-host copies of one chunk, spread over a few threads when they are large (one thread moves ~10 GB/s, the
// PCIe link 50)
void parallel_copy(const std::vector<CopyJob> &jobs) {
    size_t total = 0;
    for (const auto &j : jobs) total += j.bytes;
    static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    static const unsigned cap = []() {
        const char *v = getenv("SAF_COPY_THREADS"); // tuning knob
        // eight copy threads keep a 100 M-point pageable batch behind its kernel on a 16-thread host (terrain gradient:
        // 115 ms with six, 82.6 ms with eight against an 80.3 ms pinned run); six when several ranks share the host
        // (torchrun exports LOCAL_WORLD_SIZE), the configuration measured on the 8-GPU box
        const char *lw = getenv("LOCAL_WORLD_SIZE");
        const int n = v ? atoi(v) : (lw && atoi(lw) > 1 ? 6 : 8);
        return (unsigned)(n < 1 ? 1 : n > 16 ? 16 : n);
    }();
    const unsigned n_threads = total < ((size_t)2 << 20) ? 1u : std::min(cap, std::max(1u, hw / 2));
    if (n_threads <= 1) {
        for (const auto &j : jobs) std::memcpy(j.dst, j.src, j.bytes);
        return;
    }
    // cut every job into n_threads slices (64-byte aligned cuts)
    std::vector<std::function<void()>> tasks;
    for (unsigned t = 1; t <= n_threads; ++t) {
        tasks.emplace_back([&jobs, t, n_threads]() {
            for (const auto &j : jobs) {
                const size_t per = ((j.bytes + n_threads - 1) / n_threads + 63) & ~(size_t)63;
                const size_t b = std::min(j.bytes, per * (t - 1)), e = std::min(j.bytes, per * t);
                if (e > b) std::memcpy((char *)j.dst + b, (const char *)j.src + b, e - b);
            }
        });
    }
    g_copy_pool.run(tasks);
}

} // namespace

// Evaluates points [begin, end) of host columns on the current device.
static int eval_host_range(const saf_program *prog, const double *const *cols, const size_t *lens, size_t n_cols,
                           double *const *outs, size_t begin, size_t end, uint64_t iters, unsigned flags) {
    if (end <= begin) return SAF_OK;
    int dev = 0;
    SAF_CUDA(cudaGetDevice(&dev));
    const DeviceProgram &dp = prog->dev;
    const size_t n_in = std::min<size_t>(n_cols, dp.param_count);
    const size_t n_out = dp.result_field.size();
    const size_t total = end - begin;
    // iterated maps keep their state on chip: one chunk per stream is enough, but chunks must be
    // whole launches, so the same chunking applies.
    size_t chunk = std::min(total, kChunkPoints);
    if (total <= kChunkPoints) {
        // mid-sized batches: up to four chunks of at least 256 Ki points (enough for four points per thread on every
        // SM), so that the copies of one chunk overlap the kernel of another -- what a long in-kernel iteration
        // over one million points would otherwise wait for at both ends
        size_t c = (total + 3) / 4;
        c = (c + 1023) & ~(size_t)1023;
        chunk = std::min(total, std::max<size_t>(c, 256 * 1024));
    }
    const int n_buf = (int)std::min<size_t>(3, (total + chunk - 1) / chunk);

    // pageable columns / outputs of batches worth staging (small ones: the driver's own path is as good)
    const bool stage_ok = total * sizeof(double) >= ((size_t)1 << 20);
    std::vector<uint8_t> in_pageable(n_in, 0), out_pageable(n_out, 0);
    bool any_pageable = false;
    if (stage_ok) {
        for (size_t j = 0; j < n_in; ++j)
            if (dp.param_slot[j] != NO_SLOT && lens[j] > begin && is_pageable(cols[j])) in_pageable[j] = 1, any_pageable = true;
        for (size_t r = 0; r < n_out; ++r)
            if (is_pageable(outs[r])) out_pageable[r] = 1, any_pageable = true;
    }

    StreamSet ss;
    int rc = acquire_streams(dev, &ss);
    if (rc != SAF_OK) return rc;
    std::vector<void *> scratch, pinned;
    auto cleanup = [&]() {
        for (void *p : scratch) g_pool.release(dev, p);
        for (void *p : pinned) g_pinned.release(p);
        release_streams(dev, ss);
    };
    std::vector<std::vector<double *>> d_in(n_buf, std::vector<double *>(n_in, nullptr));
    std::vector<std::vector<double *>> d_out(n_buf, std::vector<double *>(n_out, nullptr));
    std::vector<std::vector<double *>> h_in(n_buf, std::vector<double *>(n_in, nullptr));   // pinned staging
    std::vector<std::vector<double *>> h_out(n_buf, std::vector<double *>(n_out, nullptr));
    for (int b = 0; b < n_buf; ++b) {
        for (size_t j = 0; j < n_in + n_out; ++j) {
            if (j < n_in && dp.param_slot[j] == NO_SLOT) continue; // column never read: do not copy it
            cudaError_t e = cudaSuccess;
            void *p = g_pool.acquire(dev, chunk * sizeof(double), &e);
            if (!p) {
                cleanup();
                return cuda_fail(e, "cudaMalloc(scratch)");
            }
            scratch.push_back(p);
            if (j < n_in) d_in[b][j] = (double *)p;
            else d_out[b][j - n_in] = (double *)p;
            if (j < n_in ? in_pageable[j] : out_pageable[j - n_in]) {
                void *hp = g_pinned.acquire(chunk * sizeof(double), &e);
                if (!hp) {
                    cleanup();
                    return cuda_fail(e, "cudaHostAlloc(staging)");
                }
                pinned.push_back(hp);
                if (j < n_in) h_in[b][j] = (double *)hp;
                else h_out[b][j - n_in] = (double *)hp;
            }
        }
    }
    int status = SAF_OK;
    // chunk whose results still sit in the pinned output staging of buffer b: {offset, points}
    std::vector<std::pair<size_t, size_t>> pending(n_buf, {0, 0});
    std::vector<CopyJob> jobs;
    auto drain = [&](int b) { // results of the chunk in flight on buffer b -> the caller's (pageable) outputs
        if (pending[b].second == 0) return;
        cudaError_t e = cudaStreamSynchronize(ss.s[b]);
        if (e != cudaSuccess && status == SAF_OK) status = cuda_fail(e, "cudaStreamSynchronize");
        if (status == SAF_OK) {
            jobs.clear();
            for (size_t r = 0; r < n_out; ++r)
                if (out_pageable[r]) jobs.push_back({outs[r] + pending[b].first, h_out[b][r], pending[b].second * sizeof(double)});
            parallel_copy(jobs);
        }
        pending[b] = {0, 0};
    };
    size_t k = 0;
    for (size_t off = begin; off < end && status == SAF_OK; off += chunk, ++k) {
        const int b = (int)(k % (size_t)n_buf);
        cudaStream_t st = ss.s[b];
        const size_t n = std::min(chunk, end - off);
        if (any_pageable) {
            drain(b); // also makes the staging buffers of b free to be overwritten
            if (status != SAF_OK) break;
            jobs.clear();
            for (size_t j = 0; j < n_in; ++j)
                if (in_pageable[j] && lens[j] > off)
                    jobs.push_back({h_in[b][j], cols[j] + off, std::min(n, lens[j] - off) * sizeof(double)});
            parallel_copy(jobs);
        }
        std::vector<const double *> c_ptr(n_in, nullptr);
        std::vector<size_t> c_len(n_in, 0);
        for (size_t j = 0; j < n_in; ++j) {
            if (dp.param_slot[j] == NO_SLOT) continue;
            const size_t len = lens[j];
            if (len > off) { // part (or all) of this chunk comes from the column itself
                const size_t m = std::min(n, len - off);
                const double *src = in_pageable[j] ? h_in[b][j] : cols[j] + off;
                cudaError_t e = cudaMemcpyAsync(d_in[b][j], src, m * sizeof(double), cudaMemcpyHostToDevice, st);
                if (e != cudaSuccess) { status = cuda_fail(e, "cudaMemcpyAsync(H2D)"); break; }
                c_len[j] = m;
            } else { // column exhausted before this chunk: broadcast its last element (scalar.rs:203-207)
                cudaError_t e = cudaMemcpyAsync(d_in[b][j], cols[j] + (len - 1), sizeof(double), cudaMemcpyHostToDevice, st);
                if (e != cudaSuccess) { status = cuda_fail(e, "cudaMemcpyAsync(H2D)"); break; }
                c_len[j] = 1;
            }
            c_ptr[j] = d_in[b][j];
        }
        if (status != SAF_OK) break;
        status = launch_on_device(prog, c_ptr.data(), c_len.data(), n_in, d_out[b].data(), n, iters,
                                  flags | SAF_EVAL_BROADCAST, st);
        if (status != SAF_OK) break;
        for (size_t r = 0; r < n_out; ++r) {
            double *dst = out_pageable[r] ? h_out[b][r] : outs[r] + off;
            cudaError_t e = cudaMemcpyAsync(dst, d_out[b][r], n * sizeof(double), cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) { status = cuda_fail(e, "cudaMemcpyAsync(D2H)"); break; }
        }
        if (any_pageable) pending[b] = {off, n};
    }
    // oldest chunk first: its copy-out overlaps the GPU work of the younger ones
    for (size_t i = 0; i < (size_t)n_buf; ++i) drain((int)((k + i) % (size_t)n_buf));
    for (int b = 0; b < n_buf; ++b) {
        cudaError_t e = cudaStreamSynchronize(ss.s[b]);
        if (e != cudaSuccess && status == SAF_OK) status = cuda_fail(e, "cudaStreamSynchronize");
    }
    cleanup();
    return status;
}

extern "C" int saf_eval_host(const saf_program *prog, const double *const *cols, const size_t *lens, size_t n_cols,
                             double *const *outs, size_t n_points, unsigned flags) {
    int rc = check_columns(prog, cols, lens, n_cols, outs, n_points, flags);
    if (rc != SAF_OK) return rc;
    return eval_host_range(prog, cols, lens, n_cols, outs, 0, n_points, 1, flags);
}

extern "C" int saf_iterate_host(const saf_program *prog, double *const *state, size_t n_state, size_t n_points,
                                uint64_t iters) {
    if (!prog || (n_points && !state)) return fail(SAF_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_state != prog->dev.param_count || n_state != prog->dev.result_field.size())
        return fail(SAF_ERR_INVALID_ARGUMENT, "iterate needs n_state == param_count == n_results");
    if (iters == 0 || n_points == 0) return SAF_OK;
    std::vector<size_t> lens(n_state, n_points);
    std::vector<const double *> cols(state, state + n_state);
    while (iters > 0) { // see saf_iterate_device
        const uint64_t now = std::min<uint64_t>(iters, kMaxItersPerLaunch);
        int rc = eval_host_range(prog, cols.data(), lens.data(), n_state, state, 0, n_points, now, 0);
        if (rc != SAF_OK) return rc;
        iters -= now;
    }
    return SAF_OK;
}

extern "C" int saf_eval_host_multi(const saf_program *prog, const double *const *cols, const size_t *lens,
                                   size_t n_cols, double *const *outs, size_t n_points, unsigned flags,
                                   const int *devices, int n_devices) {
    int rc = check_columns(prog, cols, lens, n_cols, outs, n_points, flags);
    if (rc != SAF_OK) return rc;
    if (n_devices <= 0) return fail(SAF_ERR_INVALID_ARGUMENT, "n_devices must be positive");
    if (n_devices == 1) {
        int prev = 0;
        SAF_CUDA(cudaGetDevice(&prev));
        SAF_CUDA(cudaSetDevice(devices ? devices[0] : 0));
        rc = eval_host_range(prog, cols, lens, n_cols, outs, 0, n_points, 1, flags);
        cudaSetDevice(prev);
        return rc;
    }
    // contiguous shards [g*N/G, (g+1)*N/G), one host thread + stream set per device, no collective
    std::vector<int> rcs(n_devices, SAF_OK);
    std::vector<std::string> errs(n_devices);
    std::vector<std::thread> th;
    for (int g = 0; g < n_devices; ++g) {
        th.emplace_back([&, g]() {
            const size_t b = (size_t)((unsigned __int128)n_points * (unsigned)g / (unsigned)n_devices);
            const size_t e = (size_t)((unsigned __int128)n_points * (unsigned)(g + 1) / (unsigned)n_devices);
            cudaError_t ce = cudaSetDevice(devices ? devices[g] : g);
            if (ce != cudaSuccess) {
                rcs[g] = SAF_ERR_CUDA;
                errs[g] = std::string("cudaSetDevice: ") + cudaGetErrorString(ce);
                return;
            }
            rcs[g] = eval_host_range(prog, cols, lens, n_cols, outs, b, e, 1, flags);
            if (rcs[g] != SAF_OK) errs[g] = g_last_error;
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < n_devices; ++g)
        if (rcs[g] != SAF_OK) return fail(rcs[g], errs[g]);
    return SAF_OK;
}

// ---- misc ---------------------------------------------------------------------------------------------------
extern "C" const char *saf_last_error(void) { return g_last_error.c_str(); }

extern "C" int saf_device_count(int *count) {
    if (!count) return fail(SAF_ERR_INVALID_ARGUMENT, "count is NULL");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return cuda_fail(e, "cudaGetDeviceCount");
    }
    return SAF_OK;
}

namespace {
// releases the probes' events and device buffers on every return path (SAF_CUDA returns early)
struct ProbeGuard {
    cudaEvent_t ev[2] = {nullptr, nullptr};
    void *mem[2] = {nullptr, nullptr};
    ~ProbeGuard() {
        for (auto e : ev) if (e) cudaEventDestroy(e);
        for (auto m : mem) if (m) cudaFree(m);
    }
};
} // namespace

extern "C" int saf_measure_fp64_peak(double *dfma_lanes_per_second) {
    if (!dfma_lanes_per_second) return fail(SAF_ERR_INVALID_ARGUMENT, "NULL argument");
    int dev = 0, sms = 0;
    SAF_CUDA(cudaGetDevice(&dev));
    SAF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ProbeGuard g;
    SAF_CUDA(cudaMalloc(&g.mem[0], 8));
    double *sink = static_cast<double *>(g.mem[0]);
    SAF_CUDA(cudaEventCreate(&g.ev[0]));
    SAF_CUDA(cudaEventCreate(&g.ev[1]));
    const cudaEvent_t e0 = g.ev[0], e1 = g.ev[1];
    const int blocks = sms * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        SAF_CUDA(cudaEventRecord(e0, 0));
        SAF_CUDA(launch_fp64_peak(sink, blocks, iters, 0));
        SAF_CUDA(cudaEventRecord(e1, 0));
        SAF_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SAF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double lanes = (double)blocks * 256.0 * (double)iters * 64.0;
        if (rep > 0 && ms > 0.f) best = std::max(best, lanes / (ms * 1e-3));
    }
    *dfma_lanes_per_second = best;
    return SAF_OK;
}

extern "C" int saf_measure_clifford_probe(size_t n_points, int iters, double *seconds) {
    if (!seconds || n_points == 0 || iters <= 0) return fail(SAF_ERR_INVALID_ARGUMENT, "bad argument");
    int dev = 0, sms = 0;
    SAF_CUDA(cudaGetDevice(&dev));
    SAF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ProbeGuard g;
    SAF_CUDA(cudaMalloc(&g.mem[0], n_points * 8));
    SAF_CUDA(cudaMalloc(&g.mem[1], n_points * 8));
    double *xs = static_cast<double *>(g.mem[0]), *ys = static_cast<double *>(g.mem[1]);
    std::vector<double> h(n_points);
    for (size_t i = 0; i < n_points; ++i) h[i] = -2.0 + 4.0 * (double)((i * 2654435761ull) % 1000003ull) / 1000003.0;
    SAF_CUDA(cudaEventCreate(&g.ev[0]));
    SAF_CUDA(cudaEventCreate(&g.ev[1]));
    const cudaEvent_t e0 = g.ev[0], e1 = g.ev[1];
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        SAF_CUDA(cudaMemcpy(xs, h.data(), n_points * 8, cudaMemcpyHostToDevice));
        SAF_CUDA(cudaMemcpy(ys, h.data(), n_points * 8, cudaMemcpyHostToDevice));
        SAF_CUDA(cudaEventRecord(e0, 0));
        SAF_CUDA(launch_clifford_probe(xs, ys, n_points, iters, sms * SAF_MINB_P4, 0));
        SAF_CUDA(cudaEventRecord(e1, 0));
        SAF_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SAF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && (best == 0.0 || ms * 1e-3 < best)) best = ms * 1e-3;
    }
    *seconds = best;
    return SAF_OK;
}

extern "C" uint64_t saf_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char *saf_version(void) { return "symbanafis-b200 0.2.0 (sm_100a)"; }
