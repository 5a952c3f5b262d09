// spec.cpp -- writes the specialised kernel of a device program (spec.h, spec_prelude.cuh).
//
// Input is the layout-independent device program that lower.cpp produced (saf_isa.h) -- the same word stream the
// interpreter's packer (pack.cpp) and the test emulator (tests/dev_emulator.py) consume.  Every instruction becomes one
// statement over SPEC_P points; the statement is the text of the interpreter's handler for that operation (kernel.cu
// hot_run / cold_op), so the arithmetic -- which intrinsic, which operand order, which leaf function -- is the same and
// so are the bits.  Slots become register arrays, constants become immediates, the accumulator becomes a local.
#include "spec.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <map>
#include <set>
#include <sstream>

#include "../../include/saf_b200.h"

namespace saf {

namespace {

struct Ins {
    uint32_t op, prefix, fd, fa, fb, fc, imm;
};

bool decode(const DeviceProgram &dp, std::vector<Ins> &out, std::string &err) {
    size_t pc = 0;
    for (;;) {
        if (pc >= dp.code.size()) {
            err = "device program has no end marker";
            return false;
        }
        const uint64_t lo = dp.code[pc++];
        const uint32_t x = (uint32_t)(lo & 0xff);
        uint32_t op, ka, kb;
        decode_xop(x, &op, &ka, &kb);
        if (op == D_END) return true;
        Ins i;
        i.op = op;
        i.prefix = (uint32_t)(lo >> 8) & 0xff;
        i.fd = (uint32_t)(lo >> 16) & 0xffff;
        i.fa = (uint32_t)(lo >> 32) & 0xffff;
        i.fb = (uint32_t)(lo >> 48) & 0xffff;
        i.fc = FIELD_NONE;
        i.imm = 0;
        if (xop_has_hi(x)) {
            if (pc >= dp.code.size()) {
                err = "device program is truncated";
                return false;
            }
            const uint64_t hi = dp.code[pc++];
            i.fc = (uint32_t)(hi & 0xffff);
            i.imm = (uint32_t)(hi >> 32);
        }
        out.push_back(i);
    }
}

// constants are named once at the top of the kernel (k<index>), bit for bit
std::string konst_literal(double v) {
    uint64_t bits = 0;
    std::memcpy(&bits, &v, 8);
    char buf[64];
    std::snprintf(buf, sizeof buf, "__longlong_as_double(0x%016llxLL)", (unsigned long long)bits);
    return buf;
}
std::string konst(uint32_t index) { return "k" + std::to_string(index); }

// value of a source field for point p: every slot is P scalar variables s<slot>_<p> -- no arrays, no loops: the
// statements are written out per point (NVVM's loop and scalar-replacement passes took 80 s on the 5000 two-trip
// loops of a 110-term gradient; straight scalar code compiles in seconds)
std::string val(uint32_t f, int p) {
    switch (field_kind(f)) {
    case KIND_S: return "s" + std::to_string(field_index(f)) + "_" + std::to_string(p);
    case KIND_K: return konst(field_index(f));
    case KIND_ACC: return "acc_" + std::to_string(p);
    default: return "dm::nan_()";
    }
}

int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

} // namespace

namespace {
#include "reroll.inc"

size_t n_trig_exp(const std::vector<Ins> &ins) {
    size_t n = 0;
    for (const Ins &i : ins)
        n += i.op == D_SIN || i.op == D_COS || i.op == D_EXP || i.op == D_EXPSQR || i.op == D_EXPSQRNEG || i.op == D_EXPNEG ||
             i.op == D_SINCOS;
    return n;
}
size_t count_instructions(const DeviceProgram &dp, size_t *n_heavy) {
    std::vector<Ins> ins;
    std::string err;
    *n_heavy = 0;
    if (!decode(dp, ins, err)) return 0;
    for (const Ins &i : ins)
        *n_heavy += i.op == D_SIN || i.op == D_COS || i.op == D_EXP || i.op == D_LN || i.op == D_EXPSQR || i.op == D_EXPSQRNEG ||
                    i.op == D_EXPNEG || i.op == D_SINCOS || i.op == D_DIV;
    return ins.size();
}
} // namespace

SpecShape spec_shape_for(const DeviceProgram &dp, bool small_batch) {
    SpecShape s;
    size_t n_heavy = 0;
    const size_t n_instr = count_instructions(dp, &n_heavy);
    const bool long_program = n_instr > 512;
    // Live values are registers now: two per slot and point, plus the temporaries of the widest leaf in flight.
    // Four points per thread while the slots leave room for it (independent dependency chains cover the 8-cycle DFMA
    // latency inside one warp, and a column tile is two 16-byte loads per thread); fewer as the live set grows, where
    // the resident warps provide the overlap instead.
    if (long_program) {
        // 100 KB of straight-line code: ONE block of 512 threads per SM that walks it in step (barriers, below), two
        // points per thread (measured on the 110-term gradient, 20 M points: 128 threads x 5 blocks, no barriers 23.6 ms;
        // 512 threads, barrier every 400 statements: P = 2 16.6 ms, 84 registers, 3.7 s of NVRTC; P = 4 15.9 ms, 128
        // registers and 9.4 s -- not worth the compile time)
        s.points_per_thread = dp.n_slots <= 28 ? 2 : 1;
        s.block = 512;
        s.min_blocks = 1;
        s.sync_every = 400;
    } else {
        s.points_per_thread = dp.n_slots <= 10 ? 4 : dp.n_slots <= 28 ? 2 : 1;
        s.block = 128;
        s.min_blocks = s.points_per_thread == 4 ? 6 : 4; // <= 80 / 128 registers per thread (single expression at 128 Mi
                                                         // points: 0.440 ms at 88 registers, 0.425 ms at 80)
        // Programs with four or more transcendental leaves bring their own independent dependency chains (speculation,
        // spec_generate: the leaves of one point interleave): two points per thread and more resident warps then beat
        // four points per thread (Clifford map, 4 leaves: P = 4 4.95 ms, P = 2 at <= 80 registers 4.81 ms)
        std::vector<Ins> ins;
        std::string err;
        if (s.points_per_thread == 4 && decode(dp, ins, err) && n_trig_exp(ins) >= 4) {
            s.points_per_thread = 2;
            s.min_blocks = 6;
        }
    }
    if (small_batch) {
        // few points: every point its own thread and a loose register cap (168): the warps that exist are all that
        // hides latency, every one of them should be resident in the first round (double pendulum, 50 k points:
        // P = 1 2.86 ms, P = 2 4.3 ms, P = 4 6.4 ms before speculation; after it 2.39 ms uncapped at 170 registers =
        // two rounds of two blocks per SM, 2.19 ms at <= 168 = three blocks per SM)
        s.points_per_thread = 1;
        s.block = 128;
        s.min_blocks = 3;
        if (long_program) s.sync_every = 400;
    }
    const int p = env_int("SAF_SPEC_P", 0);
    if (p == 1 || p == 2 || p == 4) s.points_per_thread = p;
    const int b = env_int("SAF_SPEC_BLOCK", 0);
    if (b >= 32 && b <= 1024 && b % 32 == 0) s.block = b;
    const int m = env_int("SAF_SPEC_MINB", 0);
    if (m >= 1 && m <= 16) s.min_blocks = m;
    s.sync_every = env_int("SAF_SPEC_SYNC", s.sync_every);
    // leaves with a body of tens of instructions: beyond a few dozen inlined copies they are called instead (one copy
    // per module; the pendulum's 32 call sites inline: 2.86 ms against 3.70 ms called)
    s.outline = n_heavy * (size_t)s.points_per_thread > (size_t)env_int("SAF_SPEC_INLINE_MAX", 64);
    s.speculate = env_int("SAF_SPEC_SPECULATE", 1) != 0;
    s.prefetch = env_int("SAF_SPEC_PREFETCH", 1) != 0;
    return s;
}

bool spec_batch_is_small(const SpecShape &large, unsigned long long n_points, int sms) {
    if (large.points_per_thread == 1 && large.block <= 128) return false; // the small shape would be the same kernel
    // fewer than 12 warps per SM at the large shape's points per thread: latency is hidden by warps, not by registers
    return n_points < 12ull * 32ull * (unsigned long long)large.points_per_thread * (unsigned long long)(sms > 0 ? sms : 1);
}

int spec_generate(const DeviceProgram &dp, SpecShape &shape, std::string &text, bool &can_iterate, bool &has_prefetch_entry,
                  std::string &err) {
    has_prefetch_entry = false;
    shape.rolled = false;
    std::vector<Ins> ins;
    if (!decode(dp, ins, err)) return SAF_ERR_INVALID_PROGRAM;
    const size_t max_instr = (size_t)env_int("SAF_SPEC_MAX_INSTRUCTIONS", (int)SPEC_MAX_INSTRUCTIONS); // tuning / test knob
    if (ins.size() > max_instr) {
        err = "device program has " + std::to_string(ins.size()) + " instructions; programs above " +
              std::to_string(max_instr) + " are interpreted, not specialised";
        return SAF_ERR_UNSUPPORTED;
    }
    const size_t n_params = dp.param_count, n_results = dp.result_field.size();
    can_iterate = n_results >= n_params && n_params > 0;

    // Long sums of similar terms as a loop (reroll.inc): the kernel is then a short program again -- leaves inline, no
    // barriers, 128-thread blocks, a quarter of the NVRTC time.  OPT-IN (SAF_SPEC_REROLL=1): same bits (CPU-tested), but
    // measured SLOWER than the straight-line kernel on the terrain gradient (0.565 against 0.668 of the FP64 roofline,
    // profiles/r02_j_rolled_terrain_gradient.md): one term per loop trip leaves a warp nothing to overlap its own
    // dependency chain with (stall_wait 3.5 per issue).  `shape.reroll_note` says what was rolled, or why not.
    Rolled rolled;
    std::string rolled_tables, rolled_body;
    if (ins.size() > 512 && env_int("SAF_SPEC_REROLL", 0) != 0) {
        std::string why;
        if (reroll_analyse(dp, ins, rolled, why)) {
            shape.rolled = true;
            shape.points_per_thread = env_int("SAF_SPEC_P", 2) == 4 ? 4 : env_int("SAF_SPEC_P", 2) == 1 ? 1 : 2;
            shape.block = 128;
            const int b = env_int("SAF_SPEC_BLOCK", 0);
            if (b >= 32 && b <= 1024 && b % 32 == 0) shape.block = b;
            shape.min_blocks = env_int("SAF_SPEC_MINB", 5);
            shape.sync_every = 0;
            shape.outline = false;
            std::ostringstream t, b_;
            reroll_emit_iteration(dp, rolled, shape.points_per_thread, shape.speculate, t, b_);
            rolled_tables = t.str();
            rolled_body = b_.str();
            shape.reroll_note = std::to_string(rolled.groups.size()) + " terms of " + std::to_string(rolled.classes.size()) +
                                " shapes into " + std::to_string(rolled.chains.size()) + " sums";
        } else {
            shape.reroll_note = "not rolled: " + why;
        }
        if (std::getenv("SAF_SPEC_DEBUG")) std::fprintf(stderr, "saf spec: %s\n", shape.reroll_note.c_str());
    }

    std::set<uint32_t> slots;
    auto note = [&](uint32_t f) {
        if (field_kind(f) == KIND_S) slots.insert(field_index(f));
    };
    for (const Ins &i : ins) {
        note(i.fd);
        note(i.fa);
        note(i.fb);
        note(i.fc);
    }
    for (size_t j = 0; j < n_params; ++j)
        if (dp.param_slot[j] != NO_SLOT) slots.insert(dp.param_slot[j]);
    for (uint32_t f : dp.result_field) note(f);

    const bool outline = shape.outline;
    const int P = shape.points_per_thread;
    std::ostringstream os;
    os << "// generated by spec.cpp from a device program of " << ins.size() << " instructions, " << dp.n_slots << " slots\n";
    os << "#define SPEC_P " << P << "\n#define SPEC_BLOCK " << shape.block << "\n#define SPEC_MINB " << shape.min_blocks << "\n";
    if (outline) os << "#define SPEC_OUTLINE 1\n";
    os << "#include \"spec_prelude.cuh\"\n";
    os << "using namespace saf;\n";
    size_t n_used = 0;
    for (size_t j = 0; j < n_params; ++j) n_used += dp.param_slot[j] != NO_SLOT;
    // A second entry point with next-tile prefetch for single passes (below) when the program is short and reads few
    // columns; in-kernel iteration keeps the plain one (the prefetched columns are registers that an iterated map holds
    // for nothing: Aizawa 68 -> 86 registers, 0.466 -> 0.530 ms).
    // Only where a single pass has HBM time worth hiding: >= 0.2 bytes of column traffic per FP64 slot (single expression
    // 0.44, Clifford step 0.46, Aizawa step 1.8; the pendulum's RK4 step, 0.10, is FP64 time only).
    const double bytes_per_slot = 8.0 * (double)(n_used + n_results) / (dp.fp64_slots > 1.0 ? dp.fp64_slots : 1.0);
    has_prefetch_entry = shape.prefetch && n_used > 0 && n_used * (size_t)P <= 16 && ins.size() <= 512 && bytes_per_slot >= 0.2;
    // Constants are immediates everywhere.  (Tried for long programs: a __constant__ table read through `const double
    // k17 = saf_kt[17]` -- two UMOVs per constant saved, but the compiler hoists the loads: 84 -> 128 registers, terrain
    // gradient 16.6 -> 18.8 ms.  Dropped.)
    const bool pair_leaves = env_int("SAF_SPEC_PAIR", 1) != 0;
    os << rolled_tables;
    int emit_rc = SAF_OK;
    auto emit_kernel = [&](const char *kernel_name, bool prefetch) {
    os << "extern \"C\" __global__ void __launch_bounds__(SPEC_BLOCK, SPEC_MINB) " << kernel_name
       << "(const __grid_constant__ sp::SpecDesc d) {\n";
    os << "  constexpr int P = SPEC_P;\n";
    for (size_t k = 0; k < dp.consts.size(); ++k) os << "  const double k" << k << " = " << konst_literal(dp.consts[k]) << ";\n";
    os << "  const unsigned long long TILE = (unsigned long long)SPEC_BLOCK * P, n_tiles = (d.n_points + TILE - 1) / TILE;\n";
    // Prefetch (short programs with few columns): the columns of a block's NEXT tile are loaded into registers before the
    // current tile is evaluated, so that the loads are in flight during the arithmetic -- a single pass such as
    // x^2 + sin(x) exp(-x) has as much FP64 time as HBM time per tile, and without the overlap pays their sum.
    auto emit_loads = [&](const char *base, const char *full, const char *prefix, const char *indent) {
        for (size_t j = 0; j < n_params; ++j) {
            if (dp.param_slot[j] == NO_SLOT) continue;
            os << indent << "{ double t_[P]; sp::load_col<P, SPEC_BLOCK>(d, " << j << ", " << base << ", " << full << ", t_);";
            for (int p = 0; p < P; ++p) os << " " << prefix << dp.param_slot[j] << "_" << p << " = t_[" << p << "];";
            os << " }\n";
        }
    };
    // Two tiles ahead while that costs at most 16 registers (a 1 us memory latency at 6.5 TB/s is 44 KB in flight per
    // SM; one 16-byte load per thread and tile ahead is half of that)
    const bool prefetch2 = prefetch && n_used * (size_t)P <= 4 && env_int("SAF_SPEC_PREFETCH", 1) >= 2;
    if (prefetch) {
        for (size_t j = 0; j < n_params; ++j) {
            if (dp.param_slot[j] == NO_SLOT) continue;
            os << "  double";
            for (int p = 0; p < P; ++p) os << (p ? ", " : " ") << "n" << dp.param_slot[j] << "_" << p << " = 0.0";
            if (prefetch2)
                for (int p = 0; p < P; ++p) os << ", m" << dp.param_slot[j] << "_" << p << " = 0.0";
            os << ";\n";
        }
        os << "  if (blockIdx.x < n_tiles) {\n    const unsigned long long b0_ = blockIdx.x * TILE;\n    const bool f0_ = b0_ + TILE <= d.n_points;\n";
        emit_loads("b0_", "f0_", "n", "    ");
        os << "  }\n";
        if (prefetch2) {
            os << "  if (blockIdx.x + gridDim.x < n_tiles) {\n    const unsigned long long b0_ = (blockIdx.x + gridDim.x) * TILE;\n"
                  "    const bool f0_ = b0_ + TILE <= d.n_points;\n";
            emit_loads("b0_", "f0_", "m", "    ");
            os << "  }\n";
        }
    }
    os << "  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {\n";
    os << "    const unsigned long long tile_base = tile * TILE;\n";
    os << "    const bool full = tile_base + TILE <= d.n_points;\n";
    // slots are written before they are read (validator + lowering); the initialisation only quiets the compiler
    for (int p = 0; p < P; ++p) os << "    double acc_" << p << " = 0.0, x0_" << p << " = 0.0, x1_" << p << " = 0.0;\n";
    for (uint32_t s : slots) {
        os << "    double";
        for (int p = 0; p < P; ++p) os << (p ? ", " : " ") << "s" << s << "_" << p << " = 0.0";
        os << ";\n";
    }
    if (prefetch) {
        for (size_t j = 0; j < n_params; ++j) {
            if (dp.param_slot[j] == NO_SLOT) continue;
            os << "   ";
            for (int p = 0; p < P; ++p) os << " s" << dp.param_slot[j] << "_" << p << " = n" << dp.param_slot[j] << "_" << p << ";";
            if (prefetch2)
                for (int p = 0; p < P; ++p) os << " n" << dp.param_slot[j] << "_" << p << " = m" << dp.param_slot[j] << "_" << p << ";";
            os << "\n";
        }
        const char *ahead = prefetch2 ? "2ull * gridDim.x" : "gridDim.x";
        os << "    if (tile + " << ahead << " < n_tiles) {\n      const unsigned long long b1_ = (tile + " << ahead << ") * TILE;\n"
              "      const bool f1_ = b1_ + TILE <= d.n_points;\n";
        emit_loads("b1_", "f1_", prefetch2 ? "m" : "n", "      ");
        os << "    }\n";
    } else {
        emit_loads("tile_base", "full", "s", "    ");
    }
    os << "    for (unsigned long long it = 0;;) {\n";

    // one statement per point: `tmpl` with every $a $b $c $x $y replaced by the operand / staging register of that point
    auto subst = [&](const std::string &tmpl, const Ins &i, int p) {
        std::string out;
        for (size_t k = 0; k < tmpl.size(); ++k) {
            if (tmpl[k] != '$' || k + 1 >= tmpl.size()) {
                out += tmpl[k];
                continue;
            }
            const char t = tmpl[++k];
            if (t == 'a') out += val(i.fa, p);
            else if (t == 'b') out += val(i.fb, p);
            else if (t == 'c') out += val(i.fc, p);
            else if (t == 'x') out += "x0_" + std::to_string(p);
            else if (t == 'y') out += "x1_" + std::to_string(p);
            else if (t == 'v') out += "a_.v[" + std::to_string(p) + "]";
        }
        return out;
    };
    // Long programs: a block barrier every `sync_every` statements keeps the warps of a block inside the same stretch
    // of code, so that one instruction fetch serves all of them.  A 110-term gradient is 100 KB of straight-line code
    // that each warp walks once per tile; left alone the warps drift apart, every warp streams its own copy from L2 and
    // the kernel waits for instructions (ncu: stall_no_instruction 4.3 per issue, profiles/r02_g_*).
    const int sync_every = shape.sync_every;
    const bool speculating = shape.speculate && n_trig_exp(ins) > 0 && !shape.rolled &&
                             (!shape.outline || env_int("SAF_SPEC_SPECULATE_LONG", 0) != 0);
    // fast: the speculative copy of the body -- transcendental leaves take their fast path unconditionally and only
    // record whether an argument was outside it (slow_t_ / slow_e_); see `speculate` below
    auto emit_body = [&](bool fast) {
    for (size_t n = 0; n < ins.size(); ++n) {
        const Ins &i = ins[n];
        if (sync_every > 0 && n > 0 && n % (size_t)sync_every == 0 && (fast || !speculating)) os << "      __syncthreads();\n";
        auto scalar = [&](const std::string &tmpl) {
            for (int p = 0; p < P; ++p) os << "      acc_" << p << " = " << subst(tmpl, i, p) << ";\n";
        };
        // heavy unary operand with its affine prefix -> a_.v, optionally mapped through `pre` ($v = the operand)
        auto arg_vec = [&](const std::string &pre) {
            os << "      sp::Vec<P> a_;\n";
            for (int p = 0; p < P; ++p) {
                std::string e;
                if (i.prefix == PREFIX_MUL) e = "__dmul_rn(" + val(i.fa, p) + ", " + konst(field_index(i.fb)) + ")";
                else if (i.prefix == PREFIX_FMA)
                    e = "__fma_rn(" + val(i.fa, p) + ", " + konst(field_index(i.fb)) + ", " + konst(field_index(i.fb) + 1) + ")";
                else e = val(i.fa, p);
                os << "      a_.v[" << p << "] = " << e << ";\n";
                if (!pre.empty()) os << "      a_.v[" << p << "] = " << subst(pre, i, p) << ";\n";
            }
        };
        auto leaf = [&](const char *fn_careful) {
            std::string fn = fn_careful;
            if (fast) fn = fn.substr(0, fn.size() - 1) + (fn == "exp_v" ? "f(a_, slow_e_)" : "f(a_, slow_t_)");
            else fn += "(a_)";
            os << "      { const sp::Vec<P> r_ = sp::" << fn << ";";
            for (int p = 0; p < P; ++p) os << " acc_" << p << " = r_.v[" << p << "];";
            os << " }\n";
        };
        (void)fast;
        const std::string fn = std::to_string(i.imm) + "u";
        os << "      { // [" << n << "]\n";
        bool writes_acc = true;
        switch (i.op) {
        case D_LOADK: scalar("$a"); break;
        case D_COPY: scalar("$a"); break;
        case D_NEG: scalar("-$a"); break;
        case D_SQUARE: scalar("__dmul_rn($a, $a)"); break;
        case D_CUBE: scalar("__dmul_rn(__dmul_rn($a, $a), $a)"); break;
        case D_POW4: scalar("__dmul_rn(__dmul_rn($a, $a), __dmul_rn($a, $a))"); break;
        case D_POW3_2: scalar("__dmul_rn($a, __dsqrt_rn($a))"); break;
        case D_INVPOW3_2: scalar("__ddiv_rn(1.0, __dmul_rn($a, __dsqrt_rn($a)))"); break;
        case D_INVSQRT: scalar("__ddiv_rn(1.0, __dsqrt_rn($a))"); break;
        case D_INVSQUARE: scalar("__ddiv_rn(1.0, __dmul_rn($a, $a))"); break;
        case D_INVCUBE: scalar("__ddiv_rn(1.0, __dmul_rn(__dmul_rn($a, $a), $a))"); break;
        case D_RECIP: scalar("__ddiv_rn(1.0, $a)"); break;
        case D_SQRT: scalar("__dsqrt_rn($a)"); break;
        case D_SIN: arg_vec(""); leaf("sin_v"); break;
        case D_COS: arg_vec(""); leaf("cos_v"); break;
        case D_EXP: arg_vec(""); leaf("exp_v"); break;
        case D_LN: arg_vec(""); scalar("sp::log_($v)"); break;
        case D_EXPSQR: arg_vec("__dmul_rn($v, $v)"); leaf("exp_v"); break;
        case D_EXPSQRNEG: arg_vec("-__dmul_rn($v, $v)"); leaf("exp_v"); break;
        case D_EXPNEG: arg_vec("-$v"); leaf("exp_v"); break;
        case D_RECIPEXPM1: arg_vec(""); scalar("sp::recip_expm1_($v)"); break;
        case D_SINCOS:
            // Called leaves: two adjacent, independent SinCos -- sincos(f x + p), sincos(g y + q) of one term -- share ONE
            // call over 2 P values: half the call / return / constant-load overhead, twice the chains inside the leaf.
            if (outline && !fast && pair_leaves && n + 1 < ins.size() && ins[n + 1].op == D_SINCOS && field_kind(ins[n + 1].fa) == KIND_S &&
                !(field_kind(i.fd) == KIND_S && field_index(i.fd) == field_index(ins[n + 1].fa)) &&
                !(field_kind(i.fc) == KIND_S && field_index(i.fc) == field_index(ins[n + 1].fa))) {
                const Ins &j = ins[n + 1];
                auto arg_of = [&](const Ins &q, int p) {
                    if (q.prefix == PREFIX_MUL) return "__dmul_rn(" + val(q.fa, p) + ", " + konst(field_index(q.fb)) + ")";
                    if (q.prefix == PREFIX_FMA)
                        return "__fma_rn(" + val(q.fa, p) + ", " + konst(field_index(q.fb)) + ", " + konst(field_index(q.fb) + 1) + ")";
                    return val(q.fa, p);
                };
                os << "      sp::Vec<2 * P> a_;\n";
                for (int p = 0; p < P; ++p) os << "      a_.v[" << p << "] = " << arg_of(i, p) << ";\n";
                for (int p = 0; p < P; ++p) os << "      a_.v[" << P + p << "] = " << arg_of(j, p) << ";\n";
                os << "      const sp::Vec2<2 * P> r_ = sp::sincos_v2(a_);\n";
                for (int half = 0; half < 2; ++half) {
                    const Ins &q = half ? j : i;
                    for (int p = 0; p < P; ++p) {
                        os << "      acc_" << p << " = r_.s[" << half * P + p << "];";
                        if (field_kind(q.fc) == KIND_S) os << " " << val(q.fc, p) << " = r_.c[" << half * P + p << "];";
                        if (field_kind(q.fd) == KIND_S) os << " " << val(q.fd, p) << " = acc_" << p << ";";
                        os << "\n";
                    }
                }
                os << "      } // [" << n + 1 << "] rode along\n";
                ++n;
                continue;
            }
            arg_vec("");
            os << "      const sp::Vec2<P> r_ = sp::" << (fast ? "sincos_f(a_, slow_t_)" : "sincos_v(a_)") << ";\n";
            for (int p = 0; p < P; ++p) {
                os << "      acc_" << p << " = r_.s[" << p << "];";
                if (field_kind(i.fc) == KIND_S) os << " " << val(i.fc, p) << " = r_.c[" << p << "];";
                os << "\n";
            }
            break;
        case D_BUILTIN1: scalar("sp::builtin1_(" + fn + ", $a)"); break;
        case D_BUILTIN3: scalar("sp::builtin3_(" + fn + ", $x, $y, $a)"); break;
        case D_ADD: scalar("__dadd_rn($a, $b)"); break;
        case D_SUB: scalar("__dsub_rn($a, $b)"); break;
        case D_MUL: scalar("__dmul_rn($a, $b)"); break;
        case D_NEGMUL: scalar("-__dmul_rn($a, $b)"); break;
        case D_DIV: scalar("sp::div_($a, $b)"); break;
        case D_POW: scalar("sp::pow_($a, $b)"); break;
        case D_POWI: scalar("sp::powi_($a, " + std::to_string((int32_t)i.imm) + ")"); break;
        case D_BUILTIN2: scalar("sp::builtin2_(" + fn + ", $a, $b)"); break;
        case D_BUILTIN4: scalar("sp::builtin4_(" + fn + ", $x, $y, $a, $b)"); break;
        case D_ARG2:
            // both operands are read before either staging register is written (an operand may be ACC)
            for (int p = 0; p < P; ++p)
                os << "      { const double t0_ = " << val(i.fa, p) << ", t1_ = " << val(i.fb, p) << "; x0_" << p << " = t0_; x1_" << p
                   << " = t1_; }\n";
            writes_acc = false;
            break;
        case D_FMA: scalar("__fma_rn($a, $b, $c)"); break;
        case D_FMS: scalar("__fma_rn($a, $b, -$c)"); break;
        case D_FNMA: scalar("__fma_rn(-$a, $b, $c)"); break;
        case D_FNMS: scalar("__fma_rn(-$a, $b, -$c)"); break;
        default:
            err = "device opcode " + std::to_string(i.op) + " has no specialised form";
            emit_rc = SAF_ERR_UNSUPPORTED;
            return;
        }
        if (writes_acc && field_kind(i.fd) == KIND_S)
            for (int p = 0; p < P; ++p) os << "      " << val(i.fd, p) << " = acc_" << p << ";\n";
        os << "      }\n";
    }
    };

    // Speculation (programs whose leaves are inlined): every sin / cos / sincos / exp ends in "if the argument was
    // outside the fast range, call the math library" -- a branch per leaf, i.e. a scheduling barrier per leaf: the
    // compiler cannot interleave the dependent FMA chains of two independent leaves, and a program with few points has
    // nothing else to hide their latency with (double pendulum: 24 leaves per RK4 step, 10 warps per SM).  So the body
    // is written twice.  The first copy takes every fast path unconditionally and only records the largest |argument|
    // it saw; it is ONE basic block between divisions.  If an argument was outside the fast range -- per thread, rare --
    // the parameters are restored and the second, careful copy (the leaves with their branches) recomputes the
    // iteration.  Where the fast path applies the careful leaf returns the fast path's value, so the bits are the same.
    // Long programs can speculate too (SAF_SPEC_SPECULATE_LONG=1, OFF): the fast copy has its leaves' fast paths INLINE --
    // no calls, no argument moves, no per-leaf branches -- and keeps the barriers; the careful copy keeps the called
    // leaves and has NO barriers (it runs under a per-thread condition).  Same bits (CPU-tested), but 17 s of NVRTC and,
    // at the 128 registers a 512-thread block leaves, 45.7 ms per 20 M points of the terrain gradient against 15.9.
    const bool speculate = shape.speculate && n_trig_exp(ins) > 0 && !shape.rolled &&
                           (!outline || env_int("SAF_SPEC_SPECULATE_LONG", 0) != 0);
    if (shape.rolled) {
        os << rolled_body;
    } else if (speculate) {
        for (size_t j = 0; j < n_params; ++j)
            if (dp.param_slot[j] != NO_SLOT)
                for (int p = 0; p < P; ++p) os << "      const double save" << j << "_" << p << " = s" << dp.param_slot[j] << "_" << p << ";\n";
        os << "      unsigned slow_t_ = 0u, slow_e_ = 0u;\n";
        emit_body(true);
        os << "      if (slow_t_ >= 0x41300000u || slow_e_ >= 0x4085e000u) {\n";
        for (size_t j = 0; j < n_params; ++j)
            if (dp.param_slot[j] != NO_SLOT)
                for (int p = 0; p < P; ++p) os << "      s" << dp.param_slot[j] << "_" << p << " = save" << j << "_" << p << ";\n";
        emit_body(false);
        os << "      }\n";
    } else {
        emit_body(false);
    }
    if (emit_rc != SAF_OK) return;

    os << "      if (++it >= d.iters) break;\n";
    if (can_iterate) {
        // feedback, sequential like the interpreter's (kernel.cu U_END): parameter j <- result j
        for (size_t j = 0; j < n_params; ++j) {
            if (dp.param_slot[j] == NO_SLOT) continue;
            for (int p = 0; p < P; ++p)
                os << "      s" << dp.param_slot[j] << "_" << p << " = " << val(dp.result_field[j], p) << ";\n";
        }
    }
    os << "    }\n";
    for (size_t k = 0; k < n_results; ++k) {
        os << "    { double r_[P];";
        for (int p = 0; p < P; ++p) os << " r_[" << p << "] = " << val(dp.result_field[k], p) << ";";
        os << " sp::store_out<P, SPEC_BLOCK>(d, " << k << ", tile_base, full, r_); }\n";
    }
    os << "  }\n}\n";
    };
    emit_kernel("saf_spec", false);
    if (emit_rc == SAF_OK && has_prefetch_entry) emit_kernel("saf_spec_pf", true);
    if (emit_rc != SAF_OK) return emit_rc;
    text = os.str();
    return SAF_OK;
}

} // namespace saf
