// pack_r.cpp -- device program (saf_isa.h) -> micro-instruction image of the REGISTER machine (ruop.h).
//
// The device program keeps its live values in numbered slots plus one accumulator.  Here every value gets a place
// in the interpreter's eight real registers (R0..R7, four points each) for as long as possible:
//   * values are tracked in SSA form (one value per definition); a linear scan walks the program once;
//   * operations are two-address (Rd = Rd op X): an operand that dies at the instruction lends its register to the
//     result, otherwise the first operand is copied into a free register first (RU_MOV);
//   * when no register is free the value with the furthest next use is evicted; it is stored to a shared-memory
//     slot only if no valid copy of it is there already (parameters arrive in slots), and spilled values are read
//     straight from their slot as the X operand of later operations;
//   * heavy operations have fixed registers (R1 = f(R0), sincos also R2; the fused multiply-add family, division
//     and square root work in place on R0);
//   * cold micro-ops (pow, powi, the builtins) work slot to slot outside the dispatch loop, whose registers do not
//     survive: everything live is stored in front of them;
//   * programs whose results feed back into their parameters (saf_iterate_*) keep the state in registers R7, R6, ...
//     across iterations: [load state] body [move results to the state registers] RU_END [store state] RU_EXIT.
// The same floating-point operations run on the same values in the same order as in the device program: results
// are bit-identical to the shared-memory machine's (tests/dev_emulator.py runs both images on the CPU).
#include <algorithm>
#include <climits>
#include <cstring>
#include <map>
#include <sstream>

#include "../../include/saf_b200.h"
#include "pack.h"
#include "ruop.h"

namespace saf {

namespace {

struct Dec {
    uint32_t op, prefix, fd, fa, fb, fc, imm;
};

enum { O_NONE = 0, O_VAL = 1, O_K = 2 };
struct Opnd {
    int kind = O_NONE;
    int32_t id = -1; // value id / constant index
};
struct DI { // one device instruction over values
    uint32_t op = 0, prefix = 0, imm = 0;
    Opnd s[3];
    int32_t def = -1, def2 = -1;
    uint32_t pk = 0; // constant index of the affine prefix
};
struct Val {
    std::vector<int32_t> uses; // instruction indices, ascending; NE = read at the end of the body
    int32_t reg = -1, slot = -1;
    bool slot_valid = false;
};
struct RI {
    uint32_t uop = RU_EXIT;
    int32_t d = NO_DEST;
    uint32_t a = 0, b = 0, c = 0, imm = 0, kinds = 0, cold = 0;
};
struct Loc {
    uint32_t xk, off;
};
constexpr int32_t INF = INT_MAX;

} // namespace

int pack_program_r(const DeviceProgram &dp, PackedProgram &out, std::string &err) {
    out = PackedProgram();
    constexpr int P = R_POINTS;
    out.points_per_thread = P;
    out.layout_id = 100 + P;
    const uint32_t SB = slot_bytes(P);

    // ---- decode (same word format as pack.cpp) ----
    std::vector<Dec> code;
    bool needs_staging = false;
    for (size_t i = 0; i < dp.code.size();) {
        const uint64_t lo = dp.code[i++];
        const uint32_t x = (uint32_t)(lo & 0xffu);
        uint32_t op, ka, kb;
        decode_xop(x, &op, &ka, &kb);
        Dec d{op, (uint32_t)((lo >> 8) & 0xffu), (uint32_t)((lo >> 16) & 0xffffu), (uint32_t)((lo >> 32) & 0xffffu),
              (uint32_t)(lo >> 48), FIELD_NONE, 0};
        if (xop_has_hi(x)) {
            if (i >= dp.code.size()) {
                err = "internal: truncated device program";
                return SAF_ERR_INVALID_PROGRAM;
            }
            const uint64_t hi = dp.code[i++];
            d.fc = (uint32_t)(hi & 0xffffu);
            d.imm = (uint32_t)(hi >> 32);
        }
        if (op == D_BUILTIN3 || op == D_BUILTIN4 || op == D_ARG2) needs_staging = true;
        if (op == D_END) break;
        code.push_back(d);
    }

    // ---- values ----
    std::vector<Val> V;
    std::vector<DI> I;
    auto new_val = [&]() { V.emplace_back(); return (int32_t)V.size() - 1; };
    std::vector<int32_t> cur_slot(dp.n_slots, -1), param_val(dp.param_count, -1);
    for (uint32_t j = 0; j < dp.param_count; ++j)
        if (dp.param_slot[j] != NO_SLOT) cur_slot[dp.param_slot[j]] = param_val[j] = new_val();
    {
        int32_t acc = -1;
        bool acc_is_k = false;
        uint32_t acc_k = 0;
        bool bad = false;
        auto operand = [&](uint32_t f) {
            Opnd o;
            switch (field_kind(f)) {
            case KIND_S:
                o.kind = O_VAL;
                o.id = field_index(f) < cur_slot.size() ? cur_slot[field_index(f)] : -1;
                if (o.id < 0) bad = true;
                break;
            case KIND_K: o.kind = O_K; o.id = (int32_t)field_index(f); break;
            case KIND_ACC:
                if (acc_is_k) { o.kind = O_K; o.id = (int32_t)acc_k; }
                else { o.kind = O_VAL; o.id = acc; if (acc < 0) bad = true; }
                break;
            default: break;
            }
            return o;
        };
        for (const Dec &d : code) {
            if (d.op == D_LOADK) {
                acc_is_k = true;
                acc_k = field_index(d.fa);
                continue;
            }
            DI di;
            di.op = d.op;
            di.prefix = d.prefix;
            di.imm = d.imm;
            const bool unary = (d.op >= D_FIRST_LIGHT_UNARY && d.op < D_FIRST_LIGHT_UNARY + D_N_LIGHT_UNARY) ||
                               (d.op >= D_FIRST_HEAVY_UNARY && d.op < D_FIRST_HEAVY_UNARY + D_N_HEAVY_UNARY) || d.op == D_POWI;
            const bool ternary = d.op >= D_FIRST_TERNARY && d.op < D_FIRST_TERNARY + D_N_TERNARY;
            di.s[0] = operand(d.fa);
            if (!unary) di.s[1] = operand(d.fb);
            else if (d.prefix != PREFIX_NONE) di.pk = field_index(d.fb);
            if (ternary) di.s[2] = operand(d.fc);
            if (bad) {
                err = "internal: device program reads a slot before it is written";
                return SAF_ERR_INVALID_PROGRAM;
            }
            if (d.op != D_ARG2) {
                di.def = new_val();
                if (field_kind(d.fd) == KIND_S) cur_slot[field_index(d.fd)] = di.def;
                acc = di.def;
                acc_is_k = false;
                if (d.op == D_SINCOS && field_kind(d.fc) == KIND_S) cur_slot[field_index(d.fc)] = di.def2 = new_val();
            }
            I.push_back(di);
        }
    }
    const int32_t NE = (int32_t)I.size();
    for (int32_t j = 0; j < NE; ++j)
        for (const Opnd &o : I[j].s)
            if (o.kind == O_VAL && (V[o.id].uses.empty() || V[o.id].uses.back() != j)) V[o.id].uses.push_back(j);
    const size_t n_res = dp.result_field.size();
    std::vector<Opnd> results(n_res);
    for (size_t k = 0; k < n_res; ++k) {
        const uint32_t f = dp.result_field[k];
        if (field_kind(f) == KIND_K) {
            results[k].kind = O_K;
            results[k].id = (int32_t)field_index(f);
        } else {
            results[k].kind = O_VAL;
            results[k].id = cur_slot[field_index(f)];
            if (results[k].id < 0) {
                err = "internal: result slot is never written";
                return SAF_ERR_INVALID_PROGRAM;
            }
            if (V[results[k].id].uses.empty() || V[results[k].id].uses.back() != NE) V[results[k].id].uses.push_back(NE);
        }
    }
    auto next_use = [&](int32_t v, int32_t after) -> int32_t {
        const std::vector<int32_t> &u = V[v].uses;
        auto it = std::upper_bound(u.begin(), u.end(), after);
        return it == u.end() ? INF : *it;
    };
    auto uses_left = [&](int32_t v, int32_t after) -> int {
        const std::vector<int32_t> &u = V[v].uses;
        return (int)(u.end() - std::upper_bound(u.begin(), u.end(), after));
    };

    // ---- constants ----
    std::vector<double> K(dp.consts.begin(), dp.consts.end());
    std::map<uint64_t, uint32_t> mul_pair;
    auto prefix_off = [&](const DI &di) -> uint32_t {
        if (di.prefix == PREFIX_FMA) return di.pk * 8u;
        if (di.prefix == PREFIX_MUL) {
            const double k1 = dp.consts[di.pk];
            uint64_t bits;
            std::memcpy(&bits, &k1, 8);
            auto it = mul_pair.find(bits);
            if (it == mul_pair.end()) {
                const uint32_t o = (uint32_t)K.size() * 8u;
                K.push_back(k1);
                K.push_back(-0.0); // fma(x, k1, -0.0) == x * k1 for every x
                it = mul_pair.emplace(bits, o).first;
            }
            return it->second;
        }
        return NO_PREFIX;
    };

    // ---- slots ----
    uint32_t n_slots = 0;
    std::vector<int32_t> free_slots;
    auto alloc_slot = [&]() -> int32_t {
        if (!free_slots.empty()) {
            const int32_t s = free_slots.back();
            free_slots.pop_back();
            return s;
        }
        return (int32_t)n_slots++;
    };
    out.param_off.assign(dp.param_count, -1);
    for (uint32_t j = 0; j < dp.param_count; ++j)
        if (param_val[j] >= 0) {
            Val &v = V[param_val[j]];
            v.slot = alloc_slot();
            v.slot_valid = true;
            out.param_off[j] = (int32_t)((uint32_t)v.slot * SB);
        }
    const int32_t acc_slot = alloc_slot(); // cold micro-ops always write their result here as well
    out.acc_off = (int32_t)((uint32_t)acc_slot * SB);
    if (needs_staging) {
        out.x0_off = (int32_t)((uint32_t)alloc_slot() * SB);
        out.x1_off = (int32_t)((uint32_t)alloc_slot() * SB);
    }

    // ---- register allocation + code emission ----
    std::vector<RI> U;
    int32_t reg_val[R_NREG];
    for (int r = 0; r < R_NREG; ++r) reg_val[r] = -1;
    bool failed = false;

    auto emit = [&](uint32_t uop, uint32_t a = 0, uint32_t b = 0) {
        RI u;
        u.uop = uop;
        u.a = a;
        u.b = b;
        U.push_back(u);
    };
    auto emit_st = [&](int r, int32_t slot) { emit(RU_ST + (uint32_t)r, (uint32_t)slot * SB); };
    auto spill = [&](int32_t v) { // makes the slot copy of a register-resident value valid
        Val &x = V[v];
        if (x.slot_valid) return;
        if (x.reg < 0) { failed = true; return; }
        if (x.slot < 0) x.slot = alloc_slot();
        emit_st(x.reg, x.slot);
        x.slot_valid = true;
    };
    auto loc_of = [&](const Opnd &o) -> Loc {
        if (o.kind == O_K) return Loc{XK_K, (uint32_t)o.id * 8u};
        const Val &x = V[o.id];
        if (x.reg >= 0) return Loc{(uint32_t)x.reg, 0u};
        if (!x.slot_valid) failed = true;
        return Loc{XK_S, (uint32_t)x.slot * SB};
    };
    auto emit_mov = [&](int d, const Loc &l) {
        if (l.xk == (uint32_t)d) return;
        emit(RU_MOV + (uint32_t)d * 10u + l.xk, l.off);
    };
    auto mask_of = [&](const Opnd &o) -> uint32_t { return (o.kind == O_VAL && V[o.id].reg >= 0) ? 1u << V[o.id].reg : 0u; };
    // a free register outside `exclude`; evicts the value with the furthest next use when none is free
    auto get_free = [&](uint32_t exclude, int32_t j, int prefer = -1) -> int {
        if (prefer >= 0 && !((exclude >> prefer) & 1u) && reg_val[prefer] < 0) return prefer;
        for (int r = R_NREG - 1; r >= 0; --r)
            if (!((exclude >> r) & 1u) && reg_val[r] < 0) return r;
        int best = -1;
        int64_t best_key = -1;
        for (int r = R_NREG - 1; r >= 0; --r) {
            if ((exclude >> r) & 1u) continue;
            const int32_t v = reg_val[r];
            const int32_t nu = next_use(v, j - 1);
            const int64_t key = (int64_t)nu * 2 + (V[v].slot_valid ? 1 : 0);
            if (key > best_key) { best_key = key; best = r; }
        }
        if (best < 0) { failed = true; return 0; }
        const int32_t v = reg_val[best];
        if (next_use(v, j - 1) != INF) spill(v);
        V[v].reg = -1;
        reg_val[best] = -1;
        return best;
    };
    auto any_free = [&](uint32_t exclude) -> int {
        for (int r = R_NREG - 1; r >= 0; --r)
            if (!((exclude >> r) & 1u) && reg_val[r] < 0) return r;
        return -1;
    };
    // register r must be given up: its value moves to another register (or only keeps its slot copy)
    auto vacate = [&](int r, uint32_t exclude, int32_t j) {
        const int32_t v = reg_val[r];
        if (v < 0) return;
        if (next_use(v, j - 1) != INF) {
            int f = -1;
            if (!V[v].slot_valid || uses_left(v, j - 1) > 1) f = any_free(exclude | (1u << r));
            if (f >= 0) {
                emit_mov(f, Loc{(uint32_t)r, 0u});
                reg_val[f] = v;
                V[v].reg = f;
                reg_val[r] = -1;
                return;
            }
            spill(v);
        }
        V[v].reg = -1;
        reg_val[r] = -1;
    };
    auto define = [&](int32_t v, int d) {
        const int32_t old = reg_val[d];
        if (old >= 0 && old != v) V[old].reg = -1;
        reg_val[d] = v;
        V[v].reg = d;
        V[v].slot_valid = false;
    };
    auto dies = [&](const Opnd &o, int32_t j) { return o.kind == O_VAL && next_use(o.id, j) == INF; };
    auto in_reg = [&](const Opnd &o) { return o.kind == O_VAL && V[o.id].reg >= 0; };
    // the register the next use of the value defined at j wants it in: R0 for the operand an R0-only operation
    // overwrites (fused family, division, square root ...) and for the argument of a transcendental
    auto wanted_reg = [&](int32_t v, int32_t j) -> int {
        if (v < 0) return -1;
        const int32_t u = next_use(v, j);
        if (u == INF || u >= NE) return -1;
        const DI &n = I[u];
        auto is = [&](const Opnd &o) { return o.kind == O_VAL && o.id == v; };
        if (n.op >= D_FIRST_TERNARY && n.op < D_FIRST_TERNARY + D_N_TERNARY) {
            const bool has_k = n.s[0].kind == O_K || n.s[1].kind == O_K;
            if (has_k) return (is(n.s[0]) || is(n.s[1])) ? 0 : -1; // R0 = +-(R0 * K) +- X
            return is(n.s[2]) ? 0 : -1;                             // R0 = +-(Ra * Rb) +- R0
        }
        if (n.op == D_DIV) return is(n.s[0]) ? 0 : -1;
        if (n.op == D_POW3_2 || n.op == D_INVPOW3_2 || n.op == D_INVSQRT || n.op == D_INVSQUARE || n.op == D_INVCUBE ||
            n.op == D_RECIP || n.op == D_SQRT)
            return 0;
        if (n.op == D_SIN || n.op == D_COS || n.op == D_EXP || n.op == D_LN || n.op == D_EXPSQR || n.op == D_EXPSQRNEG ||
            n.op == D_EXPNEG || n.op == D_SINCOS)
            return 0;
        return -1;
    };
    // spilled value read often from now on and a register is idle: bring it back
    auto promote = [&](const Opnd &o, uint32_t exclude, int32_t j) {
        if (o.kind != O_VAL || V[o.id].reg >= 0 || uses_left(o.id, j - 1) < 3) return;
        const int f = any_free(exclude);
        if (f < 0) return;
        emit_mov(f, loc_of(o));
        reg_val[f] = o.id;
        V[o.id].reg = f;
    };
    // the register an in-place operation on operand o may overwrite: o's own when it dies here, else a copy
    auto inplace_reg = [&](const Opnd &o, uint32_t exclude, int32_t j, int prefer) -> int {
        if (in_reg(o) && dies(o, j)) return V[o.id].reg;
        const int d = get_free(exclude | mask_of(o), j, prefer);
        emit_mov(d, loc_of(o));
        return d;
    };
    // operand o into the fixed register r (for an operation that overwrites r)
    auto to_fixed_inplace = [&](const Opnd &o, int r, uint32_t exclude, int32_t j) {
        if (o.kind == O_VAL && V[o.id].reg == r) {
            if (!dies(o, j)) vacate(r, exclude, j); // the value moves out, its bits stay in r
            return;
        }
        vacate(r, exclude | mask_of(o), j);
        emit_mov(r, loc_of(o));
    };
    auto release_dead = [&](const Opnd &o, int32_t j) {
        if (o.kind != O_VAL || next_use(o.id, j) != INF) return;
        Val &x = V[o.id];
        if (x.reg >= 0 && reg_val[x.reg] == o.id) reg_val[x.reg] = -1;
        x.reg = -1;
        if (x.slot >= 0) free_slots.push_back(x.slot);
        x.slot = -1;
        x.slot_valid = false;
    };

    // ---- iterated maps: state registers ----
    bool feedback = n_res == dp.param_count && n_res > 0 && n_res <= 5;
    std::vector<int> home(n_res, -1);
    size_t body_start = 0;
    if (feedback) {
        for (uint32_t j = 0; j < dp.param_count; ++j) {
            home[j] = R_NREG - 1 - (int)j;
            if (param_val[j] < 0) continue;
            Val &v = V[param_val[j]];
            emit_mov(home[j], Loc{XK_S, (uint32_t)v.slot * SB});
            reg_val[home[j]] = param_val[j];
            v.reg = home[j];
            // from the second iteration on the slot holds stale state: the register is the only copy
            v.slot_valid = false;
            free_slots.push_back(v.slot);
            v.slot = -1;
        }
        body_start = U.size();
    }

    static const uint32_t H_OF[] = {H1_SIN, H1_COS, H1_EXP, H1_LN, 0, H1_EXPSQR, H1_EXPSQRNEG, H1_EXPNEG, H1_SINCOS};
    for (int32_t j = 0; j < NE && !failed; ++j) {
        const DI &di = I[j];
        const Opnd &A = di.s[0], &B = di.s[1], &C = di.s[2];
        const uint32_t op = di.op;
        const uint32_t opmask = mask_of(A) | mask_of(B) | mask_of(C);
        bool cold = false;
        if (op == D_ADD || op == D_MUL || op == D_NEGMUL || op == D_SUB) {
            const bool comm = op != D_SUB;
            uint32_t rop = op == D_ADD ? RB_ADD : op == D_MUL ? RB_MUL : op == D_NEGMUL ? RB_NEGMUL : RB_SUB;
            int d;
            Opnd X;
            if (in_reg(A) && dies(A, j)) {
                d = V[A.id].reg;
                X = B;
            } else if (in_reg(B) && dies(B, j)) {
                d = V[B.id].reg;
                X = A;
                if (!comm) rop = RB_RSUB;
            } else {
                promote(B, opmask, j);
                d = get_free(opmask | mask_of(B), j, wanted_reg(di.def, j));
                emit_mov(d, loc_of(A));
                X = B;
            }
            promote(X, opmask | (1u << d), j);
            const Loc lx = loc_of(X);
            emit(RU_BIN + (rop * 8u + (uint32_t)d) * 10u + lx.xk, lx.off);
            define(di.def, d);
        } else if (op == D_COPY) {
            if (in_reg(A) && dies(A, j)) {
                define(di.def, V[A.id].reg);
            } else {
                const int d = get_free(opmask, j, wanted_reg(di.def, j));
                emit_mov(d, loc_of(A));
                define(di.def, d);
            }
        } else if (op == D_NEG || op == D_CUBE || op == D_POW4 || op == D_SQUARE) {
            const int d = inplace_reg(A, opmask, j, wanted_reg(di.def, j));
            if (op == D_SQUARE) emit(RU_BIN + (RB_MUL * 8u + (uint32_t)d) * 10u + (uint32_t)d);
            else emit(RU_UN + (op == D_NEG ? RUN_NEG : op == D_CUBE ? RUN_CUBE : RUN_POW4) * 8u + (uint32_t)d);
            define(di.def, d);
        } else if (op == D_POW3_2 || op == D_INVPOW3_2 || op == D_INVSQRT || op == D_INVSQUARE || op == D_INVCUBE ||
                   op == D_RECIP || op == D_SQRT) {
            to_fixed_inplace(A, 0, opmask, j);
            const uint32_t u = op == D_POW3_2 ? RU0_POW3_2 : op == D_INVPOW3_2 ? RU0_INVPOW3_2 : op == D_INVSQRT ? RU0_INVSQRT
                             : op == D_INVSQUARE ? RU0_INVSQUARE : op == D_INVCUBE ? RU0_INVCUBE : op == D_RECIP ? RU0_RECIP
                             : RU0_SQRT;
            emit(RU_U0 + u);
            define(di.def, 0);
        } else if (op == D_SIN || op == D_COS || op == D_EXP || op == D_LN || op == D_EXPSQR || op == D_EXPSQRNEG ||
                   op == D_EXPNEG || op == D_SINCOS) {
            // R1 = f(R0) (R2 = cos for sincos); the argument register is only read
            const uint32_t fixed = 3u | (op == D_SINCOS ? 4u : 0u);
            if (!(A.kind == O_VAL && V[A.id].reg == 0)) {
                vacate(0, opmask | fixed, j);
                emit_mov(0, loc_of(A));
                if (in_reg(A) && (V[A.id].reg == 1 || (op == D_SINCOS && V[A.id].reg == 2))) {
                    reg_val[V[A.id].reg] = -1; // the value now lives in R0
                    V[A.id].reg = 0;
                    reg_val[0] = A.id;
                }
            }
            vacate(1, opmask | fixed, j);
            if (op == D_SINCOS) vacate(2, opmask | fixed, j);
            const uint32_t pre = prefix_off(di);
            emit(RU_H0 + 2u * H_OF[op - D_SIN] + (pre != NO_PREFIX ? 1u : 0u), 0u, pre);
            define(di.def, 1);
            if (op == D_SINCOS && di.def2 >= 0) define(di.def2, 2);
        } else if (op == D_DIV) {
            if (A.kind == O_VAL && V[A.id].reg == 0 && dies(A, j)) {
                const Loc lx = loc_of(B);
                emit(RU_DIV + lx.xk, lx.off);
            } else if (B.kind == O_VAL && V[B.id].reg == 0 && dies(B, j)) {
                const Loc lx = loc_of(A);
                emit(RU_RDIV + lx.xk, lx.off);
            } else {
                to_fixed_inplace(A, 0, opmask, j);
                const Loc lx = loc_of(B);
                emit(RU_DIV + lx.xk, lx.off);
            }
            define(di.def, 0);
        } else if (op >= D_FIRST_TERNARY && op < D_FIRST_TERNARY + D_N_TERNARY) {
            // the fused family works on R0: R0 = +-(R0 * K) +- X,  R0 = +-(Ra * K) +- R0,  R0 = +-(Ra * Rb) +- R0
            const uint32_t s = op - D_FIRST_TERNARY; // FMA FMS FNMA FNMS
            Opnd M = A, Kc = B;                       // product = M * Kc with Kc constant, if any
            if (M.kind == O_K && Kc.kind != O_K) std::swap(M, Kc);
            auto same = [](const Opnd &x, const Opnd &y) { return x.kind == O_VAL && y.kind == O_VAL && x.id == y.id; };
            if (Kc.kind == O_K) {
                if (C.kind == O_VAL && V[C.id].reg == 0 && dies(C, j) && in_reg(M)) {
                    emit(RU_FKA + s * 8u + (uint32_t)V[M.id].reg, (uint32_t)Kc.id * 8u);
                } else {
                    if (same(M, C)) { // x * K +- x: the addend must survive the move of the factor into R0
                        to_fixed_inplace(M, 0, opmask, j);
                        // R0 now holds the bits of M; the value itself (if it lives on) sits elsewhere, or it dies
                        // here: read the addend from R0 itself
                        emit(RU_FKM + s * 10u + 0u, (uint32_t)Kc.id * 8u, 0u);
                    } else {
                        to_fixed_inplace(M, 0, opmask, j);
                        const Loc lc = loc_of(C);
                        emit(RU_FKM + s * 10u + lc.xk, (uint32_t)Kc.id * 8u, lc.off);
                    }
                }
            } else {
                to_fixed_inplace(C, 0, opmask, j); // the addend; a factor that shared R0 has moved (or dies here)
                uint32_t busy = 1u | mask_of(A) | mask_of(B);
                auto ensure_reg = [&](const Opnd &o) -> int {
                    if (in_reg(o)) return V[o.id].reg;
                    if (same(o, C)) return 0; // dies here, its bits are in R0
                    const int r = get_free(busy, j);
                    emit_mov(r, loc_of(o));
                    busy |= 1u << r;
                    if (o.kind == O_VAL) {
                        reg_val[r] = o.id;
                        V[o.id].reg = r;
                    }
                    return r;
                };
                const int ra = ensure_reg(A);
                busy |= 1u << ra;
                const int rb = same(A, B) ? ra : ensure_reg(B);
                emit(RU_FRR + s * R_NPAIR + r_pair((uint32_t)std::min(ra, rb), (uint32_t)std::max(ra, rb)));
            }
            define(di.def, 0);
        } else {
            cold = true;
        }
        if (cold) {
            uint32_t g;
            switch (op) {
            case D_POW: g = G_POW; break;
            case D_POWI: g = G_POWI; break;
            case D_BUILTIN1: g = G_BUILTIN1; break;
            case D_BUILTIN2: g = G_BUILTIN2; break;
            case D_BUILTIN3: g = G_BUILTIN3; break;
            case D_BUILTIN4: g = G_BUILTIN4; break;
            case D_ARG2: g = G_ARG2; break;
            case D_RECIPEXPM1: g = G_RECIPEXPM1; break;
            default: {
                std::ostringstream os;
                os << "internal: device op " << op << " has no register-machine micro-op";
                err = os.str();
                return SAF_ERR_INVALID_PROGRAM;
            }
            }
            // operands to slots, then every live register: the dispatch loop's registers do not survive
            for (const Opnd &o : di.s)
                if (o.kind == O_VAL) spill(o.id);
            for (int r = 0; r < R_NREG; ++r) {
                const int32_t v = reg_val[r];
                if (v < 0) continue;
                if (next_use(v, j) != INF) spill(v);
                V[v].reg = -1;
                reg_val[r] = -1;
            }
            RI u;
            u.uop = RU_COLD;
            u.cold = U_G + g;
            auto kind_off = [&](const Opnd &o, uint32_t &kind, uint32_t &off) {
                if (o.kind == O_K) { kind = KIND_K; off = (uint32_t)o.id * 8u; }
                else if (o.kind == O_VAL) { kind = KIND_S; off = (uint32_t)V[o.id].slot * SB; }
                else { kind = KIND_NONE; off = 0; }
            };
            uint32_t ka, kb;
            kind_off(A, ka, u.a);
            kind_off(B, kb, u.b);
            u.kinds = ka | (kb << 2) | (KIND_NONE << 4);
            u.imm = di.imm;
            if (op == D_RECIPEXPM1) u.b = prefix_off(di);
            if (di.def >= 0) {
                Val &x = V[di.def];
                x.slot = alloc_slot();
                x.slot_valid = true;
                x.reg = -1;
                u.d = (int32_t)((uint32_t)x.slot * SB);
            }
            U.push_back(u);
        }
        for (const Opnd &o : di.s) release_dead(o, j);
        if (di.def >= 0) release_dead(Opnd{O_VAL, di.def}, j);
        if (di.def2 >= 0) release_dead(Opnd{O_VAL, di.def2}, j);
    }
    if (failed) {
        err = "internal: register allocation of the register machine failed";
        return SAF_ERR_INVALID_PROGRAM;
    }

    // ---- end of the body ----
    out.result_off.assign(n_res, 0);
    out.result_is_k.assign(n_res, 0);
    if (feedback) {
        // parallel move: result k -> state register home[k]
        struct Mv { int h; Loc src; };
        std::vector<Mv> pend;
        for (size_t k = 0; k < n_res; ++k) {
            const Loc l = loc_of(results[k]);
            if (l.xk != (uint32_t)home[k]) pend.push_back(Mv{home[k], l});
        }
        while (!pend.empty()) {
            bool progress = false;
            for (size_t i = 0; i < pend.size(); ++i) {
                bool blocked = false;
                for (size_t q = 0; q < pend.size(); ++q)
                    if (q != i && pend[q].src.xk == (uint32_t)pend[i].h) blocked = true;
                if (blocked) continue;
                emit_mov(pend[i].h, pend[i].src);
                pend.erase(pend.begin() + (long)i);
                progress = true;
                break;
            }
            if (!progress) { // a cycle of registers: one source waits in a slot
                const int32_t s = alloc_slot();
                const uint32_t r = pend[0].src.xk;
                emit_st((int)r, s);
                for (Mv &m : pend)
                    if (m.src.xk == r) m.src = Loc{XK_S, (uint32_t)s * SB};
            }
        }
        RI e;
        e.uop = RU_END;
        U.push_back(e); // a = code offset of the body, filled in below
        for (size_t k = 0; k < n_res; ++k) {
            const int32_t s = alloc_slot();
            emit_st(home[k], s);
            out.result_off[k] = (int32_t)((uint32_t)s * SB);
        }
    } else {
        for (size_t k = 0; k < n_res; ++k) {
            if (results[k].kind == O_K) {
                out.result_is_k[k] = 1;
                out.result_off[k] = (int32_t)((uint32_t)results[k].id * 8u);
            } else {
                spill(results[k].id);
                out.result_off[k] = (int32_t)((uint32_t)V[results[k].id].slot * SB);
            }
        }
        if (failed) {
            err = "internal: a result of the register machine has no location";
            return SAF_ERR_INVALID_PROGRAM;
        }
    }
    {
        RI x;
        x.uop = RU_EXIT;
        U.push_back(x);
    }
    out.can_iterate = feedback;
    out.result_off_even = out.result_off;
    out.n_slots = n_slots ? n_slots : 1;
    if ((uint64_t)out.n_slots * SB >= 0x7fffffffull) {
        err = "register file offset overflow";
        return SAF_ERR_UNSUPPORTED;
    }

    // ---- windows: the last slot of every code window is RU_NEXTWIN ----
    std::vector<RI> Wd;
    std::vector<size_t> index_of(U.size());
    for (size_t i = 0; i < U.size(); ++i) {
        if (Wd.size() % CODE_WINDOW_INSTR == CODE_WINDOW_INSTR - 1) {
            RI nw;
            nw.uop = RU_NEXTWIN;
            Wd.push_back(nw);
        }
        index_of[i] = Wd.size();
        Wd.push_back(U[i]);
    }
    for (size_t i = 0; i < U.size(); ++i)
        if (U[i].uop == RU_END) Wd[index_of[i]].a = (uint32_t)(index_of[body_start] * UINSTR_BYTES);

    // ---- image: [constants, padded to 32 bytes][code] ----
    const size_t k_words = (K.size() + 3) & ~(size_t)3;
    out.image.assign(k_words + Wd.size() * UINSTR_WORDS, 0);
    for (size_t i = 0; i < K.size(); ++i) std::memcpy(&out.image[i], &K[i], 8);
    for (size_t i = 0; i < Wd.size(); ++i) {
        const RI &u = Wd[i];
        const uint32_t next = i + 1 < Wd.size() ? Wd[i + 1].uop : (uint32_t)RU_EXIT;
        uint64_t *w = &out.image[k_words + i * UINSTR_WORDS];
        w[0] = (uint64_t)next | ((uint64_t)(uint32_t)u.d << 32);
        w[1] = (uint64_t)u.a | ((uint64_t)u.b << 32);
        w[2] = (uint64_t)u.c | ((uint64_t)u.imm << 32);
        w[3] = (uint64_t)(u.kinds | (u.cold << 8)) | ((uint64_t)u.uop << 32);
    }
    out.n_instr = (uint32_t)Wd.size();
    out.code_off = (uint32_t)(k_words * 8);
    out.code_bytes = (uint32_t)(Wd.size() * UINSTR_BYTES);
    out.double_buffered = false;
    return SAF_OK;
}

} // namespace saf
