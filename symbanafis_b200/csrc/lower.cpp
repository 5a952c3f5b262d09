// lower.cpp -- reference bytecode -> device program.
//
// Pipeline (host, runs once per program handle, microseconds):
//   1. validate            bounds / def-before-use / pool ranges (helper.rs:98-196 semantics)
//   2. SSA + value numbering over ALL input programs: registers are renamed to values, Copy
//      becomes an alias, AddN/MulN/Add3/Add4/Mul3/Mul4 become left-to-right chains (same rounding
//      order as macros.rs:33-104), identical (op, operands) pairs are evaluated once, Sin(x) and
//      Cos(x) of one value share a SinCos (cf. fusion.rs:105-187).
//   3. dead-value removal from the result set
//   4. accumulator forwarding + linear-scan slot allocation: a value consumed by the very next
//      instruction is read from ACC (real registers); a value with no other consumer never gets a slot.
//   5. emit 64-bit device words + constant table.
// Nothing here changes which floating-point operations run on which values, so results are
// bit-identical to executing the reference stream instruction by instruction.
#include "lower.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>
#include <unordered_map>

#include "../../include/saf_b200.h"
#include "saf_isa.h"

namespace saf {

// operand words after each reference opcode (assemble.rs:9-103)
static const int8_t REF_OPERAND_WORDS[R_OPCOUNT] = {0, 2, 2, 3, 3, 4, 5, 3, 3, 4, 5, 3, 3, 3,
                                                    3, 4, 4, 3, 4, 4, 2, 2, 2, 2, 2, 2, 2, 2,
                                                    2, 3, 2, 2, 2, 2, 2, 2, 2, 2, 3, 4, 5, 6};

int validate(const RefProgram &p, std::string &err) {
    std::ostringstream os;
    const uint64_t ws = p.workspace_size;
    const uint64_t const_limit = (uint64_t)p.param_count + p.consts.size();
    if (const_limit > ws) {
        os << "invalid register layout: params+consts " << const_limit << " > workspace " << ws;
        err = os.str();
        return SAF_ERR_INVALID_PROGRAM;
    }
    std::vector<uint8_t> defined(ws ? ws : 1, 0);
    std::fill(defined.begin(), defined.begin() + const_limit, 1);
    size_t pc = 0;
    long idx = 0;
    auto fail = [&](const std::string &m) {
        err = m;
        return SAF_ERR_INVALID_PROGRAM;
    };
    auto rd = [&](uint32_t r) -> bool {
        if (r >= ws) {
            os << "source register R" << r << " out of bounds at instruction " << idx;
            return false;
        }
        if (!defined[r]) {
            os << "source register R" << r << " read before definition at instruction " << idx;
            return false;
        }
        return true;
    };
    auto wr = [&](uint32_t r) -> bool {
        if (r >= ws) {
            os << "destination register R" << r << " out of bounds at instruction " << idx;
            return false;
        }
        if (r >= p.param_count && r < const_limit) {
            // The reference copies constants into the register file once per batch
            // (scalar.rs:176-181, simd.rs:60-62): a program that overwrote one would behave
            // differently per point.  No compiler pass emits this; refuse it.
            os << "destination register R" << r << " is a constant-pool register at instruction "
               << idx;
            return false;
        }
        defined[r] = 1;
        return true;
    };
    for (;; ++idx) {
        if (pc >= p.flat.size()) return fail("bytecode stream not terminated by End");
        uint32_t op = p.flat[pc];
        if (op == R_END) break;
        if (op >= R_OPCOUNT) {
            os << "unknown opcode " << op << " at instruction " << idx;
            return fail(os.str());
        }
        int nw = REF_OPERAND_WORDS[op];
        if (pc + 1 + (size_t)nw > p.flat.size()) {
            os << "truncated instruction " << idx << " (opcode " << op << ")";
            return fail(os.str());
        }
        const uint32_t *w = &p.flat[pc + 1];
        bool ok = true;
        switch (op) {
        case R_SINCOS: ok = rd(w[2]) && wr(w[0]) && wr(w[1]); break;
        case R_ADDN:
        case R_MULN: {
            uint64_t start = w[1], count = w[2];
            if (count == 0 || start + count > p.pool.size()) {
                os << "arg pool slice [" << start << ".." << start + count
                   << ") out of bounds or empty at instruction " << idx;
                return fail(os.str());
            }
            for (uint64_t i = 0; ok && i < count; ++i) ok = rd(p.pool[start + i]);
            ok = ok && wr(w[0]);
        } break;
        case R_POWI: ok = rd(w[1]) && wr(w[0]); break;
        case R_BUILTIN1:
        case R_BUILTIN2:
        case R_BUILTIN3:
        case R_BUILTIN4:
            for (int i = 2; ok && i < nw; ++i) ok = rd(w[i]);
            if (ok && (w[1] & 0xffu) >= FN_COUNT) {
                os << "unknown FnOp " << (w[1] & 0xffu) << " at instruction " << idx;
                ok = false;
            }
            ok = ok && wr(w[0]);
            break;
        default:
            for (int i = 1; ok && i < nw; ++i) ok = rd(w[i]);
            ok = ok && wr(w[0]);
        }
        if (!ok) return fail(os.str());
        pc += 1 + (size_t)nw;
    }
    if (p.result_regs.empty()) return fail("program has no result register");
    for (size_t k = 0; k < p.result_regs.size(); ++k) {
        uint32_t r = p.result_regs[k];
        if (r >= ws || !defined[r]) {
            os << "result register R" << r << " undefined or out of bounds (result " << k << ")";
            return fail(os.str());
        }
    }
    return SAF_OK;
}

// ------------------------------------------------------------------------------------------------
namespace {

constexpr uint8_t V_PARAM = 250, V_CONST = 251, V_COSPART = 252;

struct Node {
    uint8_t op;
    int32_t a = -1, b = -1, c = -1, e = -1; // operand value ids
    uint32_t imm = 0;                        // FnOp id / param index
    uint64_t bits = 0;                       // constant bit pattern
};

struct Key {
    uint64_t k0, k1, k2;
    bool operator==(const Key &o) const { return k0 == o.k0 && k1 == o.k1 && k2 == o.k2; }
};
struct KeyHash {
    size_t operator()(const Key &k) const {
        uint64_t h = k.k0 * 0x9E3779B97F4A7C15ull;
        h ^= (k.k1 + 0x7F4A7C159E3779B9ull) + (h << 6) + (h >> 2);
        h ^= (k.k2 + 0x94D049BB133111EBull) + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

struct Graph {
    std::vector<Node> nodes;
    std::unordered_map<Key, int32_t, KeyHash> gvn;
    std::unordered_map<int32_t, int32_t> sincos_of; // arg value -> SINCOS node (sin part)
    std::unordered_map<int32_t, int32_t> cos_of;    // SINCOS node -> COSPART value
    uint32_t quirks = 0;

    static Key make_key(uint8_t op, uint32_t imm, int32_t a, int32_t b, int32_t c, int32_t e) {
        return Key{(uint64_t)op | ((uint64_t)imm << 8), ((uint64_t)(uint32_t)a << 32) | (uint32_t)b,
                   (uint64_t)(uint32_t)c | ((uint64_t)(uint32_t)e << 32)};
    }
    int32_t param(uint32_t i) {
        Key k{(uint64_t)V_PARAM, i, 0};
        auto it = gvn.find(k);
        if (it != gvn.end()) return it->second;
        Node n{V_PARAM};
        n.imm = i;
        nodes.push_back(n);
        return gvn[k] = (int32_t)nodes.size() - 1;
    }
    int32_t constant(double v) {
        uint64_t bits;
        std::memcpy(&bits, &v, 8);
        Key k{(uint64_t)V_CONST, bits, 0};
        auto it = gvn.find(k);
        if (it != gvn.end()) return it->second;
        Node n{V_CONST};
        n.bits = bits;
        nodes.push_back(n);
        return gvn[k] = (int32_t)nodes.size() - 1;
    }
    int32_t mk(uint8_t op, int32_t a, int32_t b = -1, int32_t c = -1, uint32_t imm = 0, int32_t e = -1) {
        // commutative normalisation (IEEE add/mul/fma are commutative in (a, b) bit for bit,
        // except for the payload of a NaN result, which the parity comparator ignores)
        if ((op == D_ADD || op == D_MUL || op == D_NEGMUL || op == D_FMA || op == D_FMS ||
             op == D_FNMA || op == D_FNMS) && a > b)
            std::swap(a, b);
        Key k = make_key(op, imm, a, b, c, e);
        auto it = gvn.find(k);
        if (it != gvn.end()) return it->second;
        Node n{op};
        n.a = a; n.b = b; n.c = c; n.e = e; n.imm = imm;
        nodes.push_back(n);
        return gvn[k] = (int32_t)nodes.size() - 1;
    }
    // Sin/Cos aware constructors: one SinCos per argument value once both halves are wanted.
    int32_t cos_part(int32_t sc) {
        auto it = cos_of.find(sc);
        if (it != cos_of.end()) return it->second;
        Node n{V_COSPART};
        n.a = sc;
        nodes.push_back(n);
        return cos_of[sc] = (int32_t)nodes.size() - 1;
    }
    int32_t sincos(int32_t x) {
        auto it = sincos_of.find(x);
        if (it != sincos_of.end()) return it->second;
        // upgrade an existing lone Sin(x) in place (it already sits before every later use)
        auto is = gvn.find(make_key(D_SIN, 0, x, -1, -1, -1));
        if (is != gvn.end()) {
            nodes[is->second].op = D_SINCOS;
            return sincos_of[x] = is->second;
        }
        Node n{D_SINCOS};
        n.a = x;
        nodes.push_back(n);
        return sincos_of[x] = (int32_t)nodes.size() - 1;
    }
    int32_t sin(int32_t x) {
        auto it = sincos_of.find(x);
        if (it != sincos_of.end()) return it->second;
        return mk(D_SIN, x);
    }
    int32_t cos(int32_t x) {
        auto it = sincos_of.find(x);
        if (it != sincos_of.end()) return cos_part(it->second);
        // a lone Sin(x) exists -> share the range reduction
        if (gvn.find(make_key(D_SIN, 0, x, -1, -1, -1)) != gvn.end()) return cos_part(sincos(x));
        return mk(D_COS, x);
    }
};

bool fn_is_unary_builtin(uint32_t fn) {
    switch (fn) {
    case FN_SIN: case FN_COS: case FN_EXP: case FN_LN: case FN_SQRT: return false; // unreachable_builtin
    default: return fn < FN_ATAN2;
    }
}
bool fn_is_binary_builtin(uint32_t fn) { return fn >= FN_ATAN2 && fn <= FN_HERMITE; }

// algorithmic FP64 issue slots per device op (SURVEY.md section 8d cost table, static SASS counts)
double op_fp64_slots(const Node &n) {
    switch (n.op) {
    case D_COPY: return 0;
    case D_NEG: case D_SQUARE: case D_ADD: case D_SUB: case D_MUL: case D_NEGMUL:
    case D_FMA: case D_FMS: case D_FNMA: case D_FNMS: return 1;
    case D_CUBE: case D_POW4: return 2;
    case D_RECIP: case D_DIV: return 10;
    case D_SQRT: return 10;
    case D_POW3_2: return 11;
    case D_INVSQRT: return 20;
    case D_INVPOW3_2: return 21;
    case D_INVSQUARE: return 11;
    case D_INVCUBE: return 12;
    case D_SIN: case D_COS: return 16;
    case D_SINCOS: return 21;
    case D_EXP: case D_EXPNEG: return 18;
    case D_EXPSQR: case D_EXPSQRNEG: return 19;
    case D_LN: return 30;
    case D_RECIPEXPM1: return 31;
    case D_POW: return 88;
    case D_POWI: return 12;
    case D_BUILTIN1:
        switch (n.imm) {
        case FN_TAN: return 31;
        case FN_SINH: case FN_COSH: return 55;
        case FN_EXPM1: return 21;
        case FN_ABS: case FN_SIGNUM: case FN_FLOOR: case FN_CEIL: case FN_ROUND: return 1;
        case FN_ERF: return 150;
        default: return 60;
        }
    case D_BUILTIN2: return n.imm == FN_ATAN2 ? 51 : 100;
    case D_BUILTIN3: case D_BUILTIN4: return 100;
    default: return 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Instruction scheduling for a small on-chip register file.
//
// The reference emits "all terms, then AddN" (sum.rs:164-237), which keeps every term of a long sum
// live at once -- harmless for an L1-resident CPU workspace, fatal for a register file that must fit
// hundreds of threads in shared memory.  The value graph fixes WHICH operations run on WHICH values
// (and therefore every rounding); the order is free.  This pass orders the live instruction nodes
//   (1) demand-driven: depth-first post-order from the results, operands left to right, so a term
//       of a left-to-right sum chain is computed immediately before the add that consumes it;
//   (2) eagerly: after every emitted instruction, any consumer of a currently live value whose
//       whole unemitted dependency cone REDUCES the number of live values is emitted at once.  This
//       interleaves the chains of different outputs (d/dx, d/dy) around their shared subexpressions
//       instead of holding those for a whole chain.
// Progress-normalised priorities.  Every result owns the interval [0, 1).  A left-deep chain of m
// adds (or muls) -- what AddN / MulN / Add3 / Add4 and the compiler's "emit all terms, then fold"
// sums become -- hands step i the sub-interval [ (i-1)/m, i/m ) of its own interval; any other node
// hands its whole interval to each operand.  A node's key is the upper end of the smallest interval
// in which it is needed (minimum over all its consumers).  Sorting by (key, creation index) is a
// valid topological order in which all long sums -- the positive and negative parts of d/dx and of
// d/dy, say -- advance in lockstep by fraction of completion, so a subexpression shared between
// term k of one sum and term k of another lives for about one term instead of a whole sum.
std::vector<int32_t> schedule_by_progress(const Graph &g, const std::vector<uint8_t> &live,
                                          const std::vector<int32_t> &results) {
    const int32_t N = (int32_t)g.nodes.size();
    auto producer = [&](int32_t v) -> int32_t {
        uint8_t op = g.nodes[v].op;
        if (op == V_COSPART) return g.nodes[v].a;
        return op < V_PARAM ? v : -1;
    };
    std::vector<int32_t> n_consumers(N, 0);
    for (int32_t n = 0; n < N; ++n) {
        if (!live[n] || g.nodes[n].op >= V_PARAM) continue;
        const Node &nd = g.nodes[n];
        for (int32_t s : {nd.a, nd.b, nd.c, nd.e})
            if (s >= 0) ++n_consumers[s];
    }
    for (int32_t r : results) ++n_consumers[r];
    std::vector<double> key(N, 2.0);
    struct Item { int32_t node; double lo, hi; };
    std::vector<Item> work;
    for (int32_t r : results) {
        int32_t pr = producer(r);
        if (pr >= 0) work.push_back({pr, 0.0, 1.0});
    }
    std::vector<int32_t> spine;
    while (!work.empty()) {
        Item it = work.back();
        work.pop_back();
        const int32_t n = it.node;
        if (it.hi >= key[n]) continue; // already needed at least this early
        key[n] = it.hi;
        const Node &nd = g.nodes[n];
        const bool chain_op = nd.op == D_ADD || nd.op == D_MUL;
        // collect the left-deep spine below n: same op, first operand used only by its parent.
        // (commutative normalisation may have swapped the operands, so either side may continue)
        spine.clear();
        if (chain_op) {
            int32_t cur = n;
            for (;;) {
                spine.push_back(cur);
                const Node &c = g.nodes[cur];
                int32_t next = -1;
                for (int32_t s : {c.a, c.b}) {
                    if (s >= 0 && g.nodes[s].op == nd.op && n_consumers[s] == 1 && producer(s) == s) {
                        next = s;
                        break;
                    }
                }
                if (next < 0) break;
                cur = next;
            }
        }
        if (spine.size() >= 3) {
            // spine[0] = top ... spine.back() = bottom; steps are numbered bottom-up
            const size_t m = spine.size() + 1; // number of terms
            const double w = (it.hi - it.lo) / (double)m;
            size_t step = 0;                   // term index, bottom-up
            for (size_t k = spine.size(); k-- > 0;) {
                const int32_t c = spine[k];
                const Node &cn = g.nodes[c];
                const int32_t below = (k + 1 < spine.size()) ? spine[k + 1] : -1;
                for (int32_t s : {cn.a, cn.b}) {
                    if (s < 0 || s == below) continue;
                    const double lo = it.lo + w * (double)step, hi = it.lo + w * (double)(step + 1);
                    ++step;
                    int32_t pr = producer(s);
                    if (pr >= 0) work.push_back({pr, lo, hi});
                }
                // the spine node itself is needed as soon as its own term is there
                if (c != n) {
                    const double hi = it.lo + w * (double)step;
                    if (hi < key[c]) key[c] = hi;
                }
            }
        } else {
            for (int32_t s : {nd.a, nd.b, nd.c, nd.e}) {
                if (s < 0) continue;
                int32_t pr = producer(s);
                if (pr >= 0) work.push_back({pr, it.lo, it.hi});
            }
        }
    }
    std::vector<int32_t> order;
    for (int32_t n = 0; n < N; ++n)
        if (live[n] && g.nodes[n].op < V_PARAM) order.push_back(n);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return key[x] < key[y]; });
    return order;
}

// Chain following on top of a given order: right after an instruction whose value has exactly one consumer, emit
// that consumer if everything else it needs has already been computed.  The value then travels through the
// accumulator instead of a register-file slot (one shared-memory store and one load less), and the consumer's
// other operands die earlier.  Any order produced this way is still topological.
std::vector<int32_t> follow_chains(const Graph &g, const std::vector<uint8_t> &live, const std::vector<int32_t> &results,
                                   const std::vector<int32_t> &base) {
    const int32_t N = (int32_t)g.nodes.size();
    auto producer = [&](int32_t v) -> int32_t {
        uint8_t op = g.nodes[v].op;
        if (op == V_COSPART) return g.nodes[v].a;
        return op < V_PARAM ? v : -1;
    };
    std::vector<int32_t> n_uses(N, 0), only_consumer(N, -1);
    for (int32_t n : base) {
        const Node &nd = g.nodes[n];
        for (int32_t s : {nd.a, nd.b, nd.c, nd.e}) {
            if (s < 0) continue;
            ++n_uses[s];
            only_consumer[s] = n;
        }
    }
    for (int32_t r : results) n_uses[r] += 2; // results stay in their slot: never chained away
    std::vector<uint8_t> two_valued(N, 0);     // SINCOS nodes produce a second value (their COSPART)
    for (auto &kv : g.cos_of) two_valued[kv.first] = 1;
    std::vector<uint8_t> done(N, 0);
    std::vector<int32_t> out;
    out.reserve(base.size());
    auto ready = [&](int32_t c, int32_t except) {
        const Node &nd = g.nodes[c];
        for (int32_t s : {nd.a, nd.b, nd.c, nd.e}) {
            if (s < 0) continue;
            const int32_t pr = producer(s);
            if (pr >= 0 && pr != except && !done[pr]) return false;
        }
        return true;
    };
    for (int32_t n0 : base) {
        int32_t n = n0;
        while (n >= 0 && !done[n]) {
            done[n] = 1;
            out.push_back(n);
            int32_t next = -1;
            if (!two_valued[n] && n_uses[n] == 1) {
                const int32_t c = only_consumer[n];
                if (c >= 0 && live[c] && !done[c] && g.nodes[c].op < V_PARAM && ready(c, n)) next = c;
            }
            n = next;
        }
    }
    return out;
}

// (c) Term-aligned merge of the long sums.  The lockstep of (a) aligns the sums by FRACTION of completion; when the
// sums of different outputs have different lengths (the positive and negative parts of d/dx and d/dy of one terrain:
// 97 / 119 / 102 / 124 terms) that drifts, and a subexpression shared by term k of two sums (sin/cos of one argument,
// one Gaussian bump) waits several terms for its second consumer.  This order advances the top-level sums ONE TERM AT A
// TIME and picks, among the next terms of all sums, the one whose emission changes the number of live values the
// least -- counting as already paid the values whose other consumer is the next term of another sum.  Two sums that
// visit their shared subexpressions in the same order (they do when they come from one symbolic expression) then
// pair up term by term and a shared value lives for one term.  Everything outside the sums is emitted on demand.
std::vector<int32_t> schedule_spine_merge(const Graph &g, const std::vector<uint8_t> &live,
                                          const std::vector<int32_t> &results) {
    const int32_t N = (int32_t)g.nodes.size();
    auto is_instr = [&](int32_t v) { return g.nodes[v].op < V_PARAM; };
    auto producer = [&](int32_t v) -> int32_t {
        uint8_t op = g.nodes[v].op;
        if (op == V_COSPART) return g.nodes[v].a;
        return op < V_PARAM ? v : -1;
    };
    auto operands = [&](int32_t n, int32_t out_vals[4]) -> int { // distinct register-file operand values
        const Node &nd = g.nodes[n];
        int k = 0;
        for (int32_t s : {nd.a, nd.b, nd.c, nd.e}) {
            if (s < 0 || producer(s) < 0) continue;
            bool dup = false;
            for (int i = 0; i < k; ++i) dup |= out_vals[i] == s;
            if (!dup) out_vals[k++] = s;
        }
        return k;
    };
    std::vector<int32_t> n_consumers(N, 0), uses_left(N, 0);
    std::vector<std::vector<int32_t>> consumers(N);
    for (int32_t n = 0; n < N; ++n) {
        if (!live[n] || !is_instr(n)) continue;
        const Node &nd = g.nodes[n];
        for (int32_t s : {nd.a, nd.b, nd.c, nd.e})
            if (s >= 0) ++n_consumers[s];
        int32_t ov[4];
        const int k = operands(n, ov);
        for (int i = 0; i < k; ++i) {
            ++uses_left[ov[i]];
            consumers[ov[i]].push_back(n);
        }
    }
    std::vector<uint8_t> is_result(N, 0);
    for (int32_t r : results) {
        ++n_consumers[r];
        is_result[r] = 1;
    }

    // ---- top-level spines: left-deep add / mul chains of >= 3 nodes met on the way down from the results ----
    struct Spine { std::vector<int32_t> nodes; size_t ptr = 0; }; // bottom-up
    std::vector<Spine> spines;
    std::vector<int32_t> spine_of(N, -1);
    {
        std::vector<uint8_t> seen(N, 0);
        std::vector<int32_t> stack;
        for (int32_t r : results) {
            const int32_t pr = producer(r);
            if (pr >= 0) stack.push_back(pr);
        }
        std::vector<int32_t> sp;
        while (!stack.empty()) {
            const int32_t n = stack.back();
            stack.pop_back();
            if (seen[n]) continue;
            seen[n] = 1;
            const Node &nd = g.nodes[n];
            if (nd.op == D_ADD || nd.op == D_MUL) {
                sp.clear();
                int32_t cur = n;
                for (;;) {
                    sp.push_back(cur);
                    const Node &c = g.nodes[cur];
                    int32_t next = -1;
                    for (int32_t s : {c.a, c.b})
                        if (s >= 0 && g.nodes[s].op == nd.op && n_consumers[s] == 1 && producer(s) == s && !seen[s]) {
                            next = s;
                            break;
                        }
                    if (next < 0) break;
                    cur = next;
                }
                if (sp.size() >= 3) {
                    Spine S;
                    S.nodes.assign(sp.rbegin(), sp.rend());
                    for (int32_t x : S.nodes) {
                        spine_of[x] = (int32_t)spines.size();
                        seen[x] = 1;
                    }
                    spines.push_back(std::move(S));
                    continue; // the terms of a spine are opaque cones: no spines are looked for inside them
                }
            }
            for (int32_t s : {nd.a, nd.b, nd.c, nd.e}) {
                if (s < 0) continue;
                const int32_t pr = producer(s);
                if (pr >= 0 && !seen[pr]) stack.push_back(pr);
            }
        }
    }

    std::vector<uint8_t> emitted(N, 0);
    std::vector<int32_t> order;
    auto emit_node = [&](int32_t n) {
        emitted[n] = 1;
        order.push_back(n);
        int32_t ov[4];
        const int k = operands(n, ov);
        for (int i = 0; i < k; ++i) --uses_left[ov[i]];
    };
    // unemitted dependency cone of an instruction, post-order
    std::vector<int32_t> stamp(N, -1), stack;
    int32_t stamp_id = 0;
    auto build_cone = [&](int32_t c, std::vector<int32_t> &cone) {
        cone.clear();
        stack.clear();
        ++stamp_id;
        stack.push_back(c);
        while (!stack.empty()) {
            int32_t x = stack.back();
            stack.pop_back();
            if (x < 0) {
                cone.push_back(~x);
                continue;
            }
            if (emitted[x] || stamp[x] == stamp_id) continue;
            stamp[x] = stamp_id;
            stack.push_back(~x);
            const Node &nd = g.nodes[x];
            const int32_t ops[4] = {nd.e, nd.c, nd.b, nd.a}; // reversed: a is visited first
            for (int32_t s : ops) {
                if (s < 0) continue;
                const int32_t pr = producer(s);
                if (pr >= 0 && !emitted[pr] && stamp[pr] != stamp_id) stack.push_back(pr);
            }
        }
    };
    std::vector<int32_t> cnt_in(N, 0), cnt_kill(N, 0), touched;
    // change in the number of live values if `cone` were emitted; `remaining` receives the values made inside the
    // cone that stay live after it
    auto cone_net = [&](const std::vector<int32_t> &cone, std::vector<int32_t> &remaining) -> int {
        int produced = 0, killed = 0;
        touched.clear();
        remaining.clear();
        for (int32_t n : cone) {
            int32_t ov[4];
            const int k = operands(n, ov);
            for (int i = 0; i < k; ++i) {
                const int32_t v = ov[i];
                if (cnt_in[v] == 0 && cnt_kill[v] == 0) touched.push_back(v);
                if (emitted[producer(v)]) ++cnt_kill[v];
                else ++cnt_in[v];
            }
        }
        for (int32_t v : touched)
            if (cnt_kill[v] > 0 && cnt_kill[v] == uses_left[v] && !is_result[v]) ++killed;
        for (int32_t n : cone) {
            auto remains = [&](int32_t v) { return is_result[v] || uses_left[v] - cnt_in[v] > 0; };
            if (remains(n)) { ++produced; remaining.push_back(n); }
            auto it = g.cos_of.find(n);
            if (it != g.cos_of.end() && live[it->second] && remains(it->second)) { ++produced; remaining.push_back(it->second); }
        }
        for (int32_t v : touched) cnt_in[v] = cnt_kill[v] = 0;
        return produced - killed;
    };

    const size_t S = spines.size();
    std::vector<std::vector<int32_t>> cones(S), remaining(S);
    std::vector<int> net(S, 0);
    std::vector<uint8_t> eligible(S, 0);
    std::vector<int32_t> in_cone_of(N, -1); // head candidate whose cone holds the node (this round)
    for (;;) {
        bool any = false;
        for (size_t s = 0; s < S; ++s) {
            Spine &sp = spines[s];
            while (sp.ptr < sp.nodes.size() && emitted[sp.nodes[sp.ptr]]) ++sp.ptr;
            cones[s].clear();
            eligible[s] = 0;
            if (sp.ptr >= sp.nodes.size()) continue;
            any = true;
            build_cone(sp.nodes[sp.ptr], cones[s]);
            eligible[s] = 1;
            for (int32_t n : cones[s])
                if (spine_of[n] >= 0 && spine_of[n] != (int32_t)s) eligible[s] = 0; // waits for another sum
            net[s] = cone_net(cones[s], remaining[s]);
        }
        if (!any) break;
        for (size_t s = 0; s < S; ++s)
            for (int32_t n : cones[s]) in_cone_of[n] = -1;
        for (size_t s = 0; s < S; ++s)
            if (eligible[s])
                for (int32_t n : cones[s])
                    if (in_cone_of[n] < 0) in_cone_of[n] = (int32_t)s; // shared nodes: first head wins
        int best = -1;
        double best_key = 0.0, best_progress = 0.0;
        for (size_t s = 0; s < S; ++s) {
            if (cones[s].empty()) continue;
            // values this term leaves behind whose (other) consumer is the next term of another sum are about to be
            // consumed: they do not count against it
            int paired = 0;
            for (int32_t v : remaining[s]) {
                if (is_result[v]) continue;
                bool hit = false;
                for (int32_t c : consumers[v])
                    if (!emitted[c] && in_cone_of[c] >= 0 && in_cone_of[c] != (int32_t)s) hit = true;
                paired += hit;
            }
            const double progress = (double)spines[s].ptr / (double)spines[s].nodes.size();
            // ineligible candidates (their term needs another sum's result) only run when nothing else can
            const double key = (eligible[s] ? 0.0 : 1e6) + (double)(net[s] - paired);
            if (best < 0 || key < best_key || (key == best_key && progress < best_progress)) {
                best = (int)s;
                best_key = key;
                best_progress = progress;
            }
        }
        const std::vector<int32_t> todo = cones[best];
        for (int32_t n : todo)
            if (!emitted[n]) emit_node(n);
    }
    // ---- everything else, on demand from the results ----
    std::vector<int32_t> cone;
    for (int32_t r : results) {
        const int32_t pr = producer(r);
        if (pr < 0 || emitted[pr]) continue;
        build_cone(pr, cone);
        const std::vector<int32_t> todo = cone;
        for (int32_t n : todo)
            if (!emitted[n]) emit_node(n);
    }
    return order;
}

std::vector<int32_t> schedule_nodes(const Graph &g, const std::vector<uint8_t> &live,
                                    const std::vector<int32_t> &results) {
    const int32_t N = (int32_t)g.nodes.size();
    auto is_instr = [&](int32_t v) { return g.nodes[v].op < V_PARAM; };
    auto producer = [&](int32_t v) -> int32_t { // instruction node that makes value v, -1 for params/consts
        uint8_t op = g.nodes[v].op;
        if (op == V_COSPART) return g.nodes[v].a;
        return op < V_PARAM ? v : -1;
    };
    // operand values of an instruction node (register-file values only)
    auto operands = [&](int32_t n, int32_t out_vals[4]) -> int {
        const Node &nd = g.nodes[n];
        int k = 0;
        for (int32_t s : {nd.a, nd.b, nd.c, nd.e}) {
            if (s < 0 || producer(s) < 0) continue;
            bool dup = false;
            for (int i = 0; i < k; ++i) dup |= out_vals[i] == s;
            if (!dup) out_vals[k++] = s;
        }
        return k;
    };
    std::vector<std::vector<int32_t>> consumers(N);
    std::vector<int32_t> uses_left(N, 0);
    std::vector<uint8_t> is_result(N, 0);
    std::vector<int32_t> instrs;
    for (int32_t n = 0; n < N; ++n) {
        if (!live[n] || !is_instr(n)) continue;
        instrs.push_back(n);
        int32_t ov[4];
        int k = operands(n, ov);
        for (int i = 0; i < k; ++i) {
            consumers[ov[i]].push_back(n);
            ++uses_left[ov[i]];
        }
    }
    for (int32_t r : results) is_result[r] = 1;

    std::vector<uint8_t> emitted(N, 0);
    std::vector<int32_t> order;
    order.reserve(instrs.size());
    std::vector<int32_t> live_vals; // values produced, still awaiting consumers
    std::vector<uint8_t> in_live(N, 0);
    std::vector<int32_t> stamp(N, -1), cnt_in(N, 0), cnt_kill(N, 0);
    int32_t stamp_id = 0;
    int32_t last_emitted = -1;

    auto emit_node = [&](int32_t n) {
        emitted[n] = 1;
        order.push_back(n);
        last_emitted = n;
        int32_t ov[4];
        int k = operands(n, ov);
        for (int i = 0; i < k; ++i) --uses_left[ov[i]];
        auto add_live = [&](int32_t v) {
            if ((uses_left[v] > 0) && !in_live[v]) {
                in_live[v] = 1;
                live_vals.push_back(v);
            }
        };
        add_live(n);
        auto it = g.cos_of.find(n);
        if (it != g.cos_of.end() && live[it->second]) add_live(it->second);
    };
    // unemitted dependency cone of instruction c in post-order; false when larger than `bound`
    std::vector<int32_t> cone, stack;
    auto build_cone = [&](int32_t c, size_t bound) -> bool {
        cone.clear();
        stack.clear();
        ++stamp_id;
        stack.push_back(c);
        // iterative post-order: a node is re-pushed negated once its operands are scheduled
        while (!stack.empty()) {
            int32_t x = stack.back();
            stack.pop_back();
            if (x < 0) {
                cone.push_back(~x);
                if (cone.size() > bound) return false;
                continue;
            }
            if (emitted[x] || stamp[x] == stamp_id) continue;
            stamp[x] = stamp_id;
            stack.push_back(~x);
            const Node &nd = g.nodes[x];
            int32_t ops[4] = {nd.e, nd.c, nd.b, nd.a}; // reversed: a is visited first
            for (int32_t s : ops) {
                if (s < 0) continue;
                int32_t pr = producer(s);
                if (pr >= 0 && !emitted[pr] && stamp[pr] != stamp_id) stack.push_back(pr);
            }
        }
        return true;
    };
    // change in the number of live values if the cone currently in `cone` were emitted
    auto cone_net = [&]() -> int {
        int produced = 0, killed = 0;
        std::vector<int32_t> touched;
        for (int32_t n : cone) {
            int32_t ov[4];
            int k = operands(n, ov);
            for (int i = 0; i < k; ++i) {
                int32_t v = ov[i];
                if (cnt_in[v] == 0 && cnt_kill[v] == 0) touched.push_back(v);
                if (emitted[producer(v)]) ++cnt_kill[v]; // consumer of an already live value
                else ++cnt_in[v];                         // consumer of a value made inside the cone
            }
        }
        for (int32_t v : touched) {
            if (cnt_kill[v] > 0 && cnt_kill[v] == uses_left[v] && !is_result[v]) ++killed;
        }
        for (int32_t n : cone) {
            auto remains = [&](int32_t v) { return is_result[v] || uses_left[v] - cnt_in[v] > 0; };
            if (remains(n)) ++produced;
            auto it = g.cos_of.find(n);
            if (it != g.cos_of.end() && live[it->second] && remains(it->second)) ++produced;
        }
        for (int32_t v : touched) cnt_in[v] = cnt_kill[v] = 0;
        return produced - killed;
    };
    auto eager = [&]() {
        for (;;) {
            // compact the live list
            size_t w = 0;
            for (int32_t v : live_vals) {
                if (uses_left[v] > 0) live_vals[w++] = v;
                else in_live[v] = 0;
            }
            live_vals.resize(w);
            int best_net = 0, best_c = -1;
            size_t best_size = 0;
            for (int32_t v : live_vals) {
                for (int32_t c : consumers[v]) {
                    if (emitted[c]) continue;
                    if (!build_cone(c, 48)) continue;
                    int net = cone_net();
                    bool uses_last = false;
                    if (cone.size() == 1) {
                        const Node &nd = g.nodes[c];
                        uses_last = nd.a == last_emitted || nd.b == last_emitted || nd.c == last_emitted;
                    }
                    bool better = net < best_net || (net == best_net && best_c >= 0 &&
                                  (cone.size() < best_size || (cone.size() == best_size && uses_last)));
                    if (net < 0 && (best_c < 0 || better)) {
                        best_net = net;
                        best_c = c;
                        best_size = cone.size();
                    }
                }
            }
            if (best_c < 0) return;
            build_cone(best_c, (size_t)-1);
            std::vector<int32_t> todo = cone;
            for (int32_t n : todo) emit_node(n);
        }
    };

    // demand-driven outer loop over the results
    std::vector<std::pair<int32_t, int>> dfs; // (node, next operand index)
    for (int32_t r : results) {
        int32_t root = producer(r);
        if (root < 0 || emitted[root]) continue;
        dfs.clear();
        dfs.push_back({root, 0});
        while (!dfs.empty()) {
            auto &top = dfs.back();
            int32_t n = top.first;
            if (emitted[n]) {
                dfs.pop_back();
                continue;
            }
            const Node &nd = g.nodes[n];
            int32_t ops[4] = {nd.a, nd.b, nd.c, nd.e};
            bool descended = false;
            while (top.second < 4) {
                int32_t s = ops[top.second++];
                if (s < 0) continue;
                int32_t pr = producer(s);
                if (pr >= 0 && !emitted[pr]) {
                    dfs.push_back({pr, 0});
                    descended = true;
                    break;
                }
            }
            if (descended) continue;
            dfs.pop_back();
            if (!emitted[n]) {
                emit_node(n);
                eager();
            }
        }
    }
    return order;
}

} // namespace

int lower(const std::vector<const RefProgram *> &progs, DeviceProgram &out, std::string &err) {
    if (progs.empty()) {
        err = "no program to lower";
        return SAF_ERR_INVALID_ARGUMENT;
    }
    Graph g;
    std::vector<int32_t> results;
    uint32_t param_count = 0;
    out = DeviceProgram();
    const double NaN = std::numeric_limits<double>::quiet_NaN();

    for (const RefProgram *pp : progs) {
        const RefProgram &p = *pp;
        int rc = validate(p, err);
        if (rc != SAF_OK) return rc;
        param_count = std::max(param_count, p.param_count);
        std::vector<int32_t> R(p.workspace_size ? p.workspace_size : 1, -1);
        for (uint32_t i = 0; i < p.param_count; ++i) R[i] = g.param(i);
        for (size_t i = 0; i < p.consts.size(); ++i) R[p.param_count + i] = g.constant(p.consts[i]);
        size_t pc = 0;
        for (;;) {
            uint32_t op = p.flat[pc];
            if (op == R_END) break;
            const uint32_t *w = &p.flat[pc + 1];
            pc += 1 + (size_t)REF_OPERAND_WORDS[op];
            ++out.n_ref_instructions;
            switch (op) {
            case R_COPY: R[w[0]] = R[w[1]]; break;
            case R_NEG: R[w[0]] = g.mk(D_NEG, R[w[1]]); break;
            case R_SINCOS: {
                int32_t sc = g.sincos(R[w[2]]);
                int32_t cs = g.cos_part(sc);
                R[w[0]] = sc;
                R[w[1]] = cs; // written last: wins when both destinations coincide (macros.rs:30-31)
            } break;
            case R_ADD: R[w[0]] = g.mk(D_ADD, R[w[1]], R[w[2]]); break;
            case R_ADD3: R[w[0]] = g.mk(D_ADD, g.mk(D_ADD, R[w[1]], R[w[2]]), R[w[3]]); break;
            case R_ADD4:
                R[w[0]] = g.mk(D_ADD, g.mk(D_ADD, g.mk(D_ADD, R[w[1]], R[w[2]]), R[w[3]]), R[w[4]]);
                break;
            case R_ADDN:
            case R_MULN: {
                uint8_t dop = op == R_ADDN ? D_ADD : D_MUL;
                int32_t acc = R[p.pool[w[1]]];
                for (uint32_t i = 1; i < w[2]; ++i) acc = g.mk(dop, acc, R[p.pool[w[1] + i]]);
                R[w[0]] = acc;
            } break;
            case R_MUL: R[w[0]] = g.mk(D_MUL, R[w[1]], R[w[2]]); break;
            case R_MUL3: R[w[0]] = g.mk(D_MUL, g.mk(D_MUL, R[w[1]], R[w[2]]), R[w[3]]); break;
            case R_MUL4:
                R[w[0]] = g.mk(D_MUL, g.mk(D_MUL, g.mk(D_MUL, R[w[1]], R[w[2]]), R[w[3]]), R[w[4]]);
                break;
            case R_SUB: R[w[0]] = g.mk(D_SUB, R[w[1]], R[w[2]]); break;
            case R_DIV: R[w[0]] = g.mk(D_DIV, R[w[1]], R[w[2]]); break;
            case R_POW: R[w[0]] = g.mk(D_POW, R[w[1]], R[w[2]]); break;
            case R_MULADD: R[w[0]] = g.mk(D_FMA, R[w[1]], R[w[2]], R[w[3]]); break;
            case R_MULSUB: R[w[0]] = g.mk(D_FMS, R[w[1]], R[w[2]], R[w[3]]); break;
            case R_NEGMUL: R[w[0]] = g.mk(D_NEGMUL, R[w[1]], R[w[2]]); break;
            case R_NEGMULADD: R[w[0]] = g.mk(D_FNMA, R[w[1]], R[w[2]], R[w[3]]); break;
            case R_NEGMULSUB: R[w[0]] = g.mk(D_FNMS, R[w[1]], R[w[2]], R[w[3]]); break;
            case R_SQUARE: R[w[0]] = g.mk(D_SQUARE, R[w[1]]); break;
            case R_CUBE: R[w[0]] = g.mk(D_CUBE, R[w[1]]); break;
            case R_POW4: R[w[0]] = g.mk(D_POW4, R[w[1]]); break;
            case R_POW3_2: R[w[0]] = g.mk(D_POW3_2, R[w[1]]); break;
            case R_INVPOW3_2: R[w[0]] = g.mk(D_INVPOW3_2, R[w[1]]); break;
            case R_INVSQRT: R[w[0]] = g.mk(D_INVSQRT, R[w[1]]); break;
            case R_INVSQUARE: R[w[0]] = g.mk(D_INVSQUARE, R[w[1]]); break;
            case R_INVCUBE: R[w[0]] = g.mk(D_INVCUBE, R[w[1]]); break;
            case R_RECIP: R[w[0]] = g.mk(D_RECIP, R[w[1]]); break;
            case R_POWI: R[w[0]] = g.mk(D_POWI, R[w[1]], g.constant((double)(int32_t)w[2])); break;
            case R_SIN: R[w[0]] = g.sin(R[w[1]]); break;
            case R_COS: R[w[0]] = g.cos(R[w[1]]); break;
            case R_EXP: R[w[0]] = g.mk(D_EXP, R[w[1]]); break;
            case R_LN: R[w[0]] = g.mk(D_LN, R[w[1]]); break;
            case R_SQRT: R[w[0]] = g.mk(D_SQRT, R[w[1]]); break;
            case R_RECIPEXPM1: R[w[0]] = g.mk(D_RECIPEXPM1, R[w[1]]); break;
            case R_EXPSQR: R[w[0]] = g.mk(D_EXPSQR, R[w[1]]); break;
            case R_EXPSQRNEG: R[w[0]] = g.mk(D_EXPSQRNEG, R[w[1]]); break;
            case R_BUILTIN1: {
                uint32_t fn = w[1] & 0xffu; // transmute_copy reads the low byte (macros.rs:311)
                if (fn == FN_EXPNEG) R[w[0]] = g.mk(D_EXPNEG, R[w[2]]);
                else if (fn_is_unary_builtin(fn)) R[w[0]] = g.mk(D_BUILTIN1, R[w[2]], -1, -1, fn);
                else { // unreachable_builtin -> NaN in the reference's release build (builtins.rs:21-26)
                    R[w[0]] = g.constant(NaN);
                    g.quirks |= SAF_QUIRK_BUILTIN1_UNREACHABLE;
                }
            } break;
            case R_BUILTIN2: {
                uint32_t fn = w[1] & 0xffu;
                if (fn_is_binary_builtin(fn)) R[w[0]] = g.mk(D_BUILTIN2, R[w[2]], R[w[3]], -1, fn);
                else R[w[0]] = g.constant(NaN);
            } break;
            case R_BUILTIN3: {
                uint32_t fn = w[1] & 0xffu;
                if (fn == FN_ASSOCLEGENDRE) R[w[0]] = g.mk(D_BUILTIN3, R[w[2]], R[w[3]], R[w[4]], fn);
                else R[w[0]] = g.constant(NaN);
            } break;
            case R_BUILTIN4: {
                uint32_t fn = w[1] & 0xffu;
                if (fn == FN_SPHERICALHARMONIC)
                    R[w[0]] = g.mk(D_BUILTIN4, R[w[2]], R[w[3]], R[w[4]], fn, R[w[5]]);
                else R[w[0]] = g.constant(NaN);
            } break;
            default: err = "internal: unhandled opcode"; return SAF_ERR_INVALID_PROGRAM;
            }
        }
        for (uint32_t r : p.result_regs) results.push_back(R[r]);
    }

    // results that are parameters get their own value so that result slots never alias parameter
    // slots (the in-kernel iteration copies results back over the parameters)
    for (int32_t &r : results)
        if (g.nodes[r].op == V_PARAM) r = g.mk(D_COPY, r);

    const int32_t N = (int32_t)g.nodes.size();
    // ---- liveness ----
    std::vector<uint8_t> live(N, 0);
    for (int32_t r : results) live[r] = 1;
    for (int32_t i = N - 1; i >= 0; --i) {
        if (!live[i]) continue;
        const Node &n = g.nodes[i];
        if (n.op == V_PARAM || n.op == V_CONST) continue;
        for (int32_t s : {n.a, n.b, n.c, n.e})
            if (s >= 0) live[s] = 1;
    }
    // a COSPART may be created after its SINCOS node; the loop above visits higher ids first, and
    // COSPART ids are always higher than their SINCOS node, so the SINCOS is marked in time.
    // A SINCOS whose cos half is dead and sin half is live (or vice versa) still runs as one op.
    for (auto &kv : g.cos_of)
        if (live[kv.second]) live[kv.first] = 1;

    // ---- emit a device program for a given instruction order ----
    const uint32_t n_ref_instructions = out.n_ref_instructions;
    auto emit_program = [&](const std::vector<int32_t> &order, DeviceProgram &out) -> int {
    out = DeviceProgram();
    out.n_ref_instructions = n_ref_instructions;

    struct Emit {
        int32_t node;     // producing node, -1 for ARG2 / LOADK
        uint8_t op;
        int32_t src[3];   // value ids
        int32_t cosval;   // for SINCOS
        uint8_t prefix = PREFIX_NONE; // affine prefix folded into a heavy unary op
        int32_t pk1 = -1, pk2 = -1;   // constant nodes of the prefix
    };
    std::vector<Emit> E;
    constexpr int32_t SRC_LOADK = -2; // operand comes from ACC, loaded by a D_LOADK emitted just before
    auto is_const = [&](int32_t v) { return v >= 0 && g.nodes[v].op == V_CONST; };
    for (int32_t i : order) {
        const Node &n = g.nodes[i];
        if (n.op == D_BUILTIN3) {
            E.push_back({-1, (uint8_t)D_ARG2, {n.a, n.b, -1}, -1});
            E.push_back({i, n.op, {n.c, -1, -1}, -1});
        } else if (n.op == D_BUILTIN4) {
            E.push_back({-1, (uint8_t)D_ARG2, {n.a, n.b, -1}, -1});
            E.push_back({i, n.op, {n.c, n.e, -1}, -1});
        } else if (n.op == D_SINCOS) {
            auto it = g.cos_of.find(i);
            int32_t cv = (it != g.cos_of.end() && live[it->second]) ? it->second : -1;
            E.push_back({i, n.op, {n.a, -1, -1}, cv});
        } else if (n.op == D_POWI) {
            E.push_back({i, n.op, {n.a, -1, -1}, -1}); // the exponent travels as an immediate
        } else if (n.op >= D_FIRST_LIGHT_UNARY && n.op < D_FIRST_LIGHT_UNARY + D_N_LIGHT_UNARY && is_const(n.a)) {
            // the kind-specialised light opcodes have no constant-operand form for these shapes:
            // materialise the constant in ACC first
            E.push_back({-1, (uint8_t)D_LOADK, {n.a, -1, -1}, -1});
            E.push_back({i, n.op, {SRC_LOADK, -1, -1}, -1});
        } else if (n.op >= D_FIRST_LIGHT_BINARY && n.op < D_FIRST_LIGHT_BINARY + D_N_LIGHT_BINARY &&
                   is_const(n.a) && is_const(n.b)) {
            E.push_back({-1, (uint8_t)D_LOADK, {n.a, -1, -1}, -1});
            E.push_back({i, n.op, {SRC_LOADK, n.b, -1}, -1});
        } else {
            E.push_back({i, n.op, {n.a, n.b, n.c}, -1});
        }
        out.fp64_slots += op_fp64_slots(n);
    }

    // ---- affine-prefix fusion: f(Mul(x, K)) and f(MulAdd(x, K1, K2)) become one instruction ----
    {
        std::vector<int32_t> uses(N, 0);
        for (const Emit &e : E)
            for (int32_t s_ : e.src)
                if (s_ >= 0) ++uses[s_];
        for (int32_t r : results) ++uses[r];
        for (auto &kv : g.cos_of)
            if (live[kv.second]) uses[kv.first] += 0; // the cos half is a separate value id
        auto prefixable = [](uint8_t op) {
            return op == D_SIN || op == D_COS || op == D_EXP || op == D_LN || op == D_EXPNEG || op == D_SINCOS ||
                   op == D_RECIPEXPM1 || op == D_EXPSQR || op == D_EXPSQRNEG;
        };
        std::vector<Emit> F;
        F.reserve(E.size());
        for (size_t j = 0; j < E.size(); ++j) {
            Emit e = E[j];
            if (!F.empty() && prefixable(e.op) && e.src[0] >= 0 && F.back().node == e.src[0] &&
                F.back().node >= 0 && uses[e.src[0]] == 1 && F.back().prefix == PREFIX_NONE) {
                const Emit &m = F.back();
                const Node &mn = g.nodes[m.node];
                int32_t x = -1, k1 = -1, k2 = -1;
                if (m.op == D_MUL) {
                    if (is_const(mn.a) && !is_const(mn.b)) { x = mn.b; k1 = mn.a; }
                    else if (is_const(mn.b) && !is_const(mn.a)) { x = mn.a; k1 = mn.b; }
                } else if (m.op == D_FMA && is_const(mn.c)) {
                    if (is_const(mn.a) && !is_const(mn.b)) { x = mn.b; k1 = mn.a; k2 = mn.c; }
                    else if (is_const(mn.b) && !is_const(mn.a)) { x = mn.a; k1 = mn.b; k2 = mn.c; }
                }
                // the producer must really have been emitted as that plain Mul / MulAdd
                if (x >= 0 && m.src[0] != SRC_LOADK) {
                    e.src[0] = x;
                    e.prefix = (k2 >= 0) ? PREFIX_FMA : PREFIX_MUL;
                    e.pk1 = k1;
                    e.pk2 = k2;
                    F.pop_back();
                }
            }
            F.push_back(e);
        }
        E.swap(F);
    }

    // ---- accumulator forwarding + last uses ----
    const int32_t NE = (int32_t)E.size();
    const int32_t INF = std::numeric_limits<int32_t>::max();
    std::vector<int32_t> last_use(N, -1);
    std::vector<uint8_t> needs_slot(N, 0);
    auto is_instr_value = [&](int32_t v) { return g.nodes[v].op < V_PARAM; };
    for (int32_t j = 0; j < NE; ++j) {
        for (int32_t s : E[j].src) {
            if (s < 0 || g.nodes[s].op == V_CONST) continue;
            bool fwd = j > 0 && E[j - 1].node == s && is_instr_value(s);
            if (fwd) {
                ++out.n_acc_forwards;
            } else {
                needs_slot[s] = 1;
                last_use[s] = j;
            }
        }
    }
    for (int32_t r : results)
        if (g.nodes[r].op != V_CONST) {
            needs_slot[r] = 1;
            last_use[r] = INF;
        }

    // ---- slots: parameters pinned first, temporaries by linear scan ----
    std::vector<int32_t> slot(N, -1);
    uint32_t n_slots = 0;
    out.param_count = param_count;
    out.param_slot.assign(param_count, NO_SLOT);
    for (int32_t i = 0; i < N; ++i)
        if (g.nodes[i].op == V_PARAM && needs_slot[i]) {
            slot[i] = (int32_t)n_slots;
            out.param_slot[g.nodes[i].imm] = (uint16_t)n_slots++;
        }
    std::vector<int32_t> free_slots;
    std::vector<std::vector<int32_t>> dying(NE);
    for (int32_t v = 0; v < N; ++v)
        if (needs_slot[v] && g.nodes[v].op != V_PARAM && last_use[v] >= 0 && last_use[v] != INF)
            dying[last_use[v]].push_back(v);
    auto alloc = [&]() -> int32_t {
        if (!free_slots.empty()) {
            int32_t s = free_slots.back();
            free_slots.pop_back();
            return s;
        }
        return (int32_t)n_slots++;
    };

    // ---- constants ----
    std::unordered_map<uint64_t, uint32_t> kidx;
    auto kfield = [&](int32_t v) -> uint32_t {
        uint64_t bits = g.nodes[v].bits;
        auto it = kidx.find(bits);
        uint32_t i;
        if (it == kidx.end()) {
            i = (uint32_t)out.consts.size();
            double d;
            std::memcpy(&d, &bits, 8);
            out.consts.push_back(d);
            kidx[bits] = i;
        } else
            i = it->second;
        return make_field(KIND_K, i & (MAX_INDEX - 1));
    };

    // (K1, K2) pairs of an FMA prefix sit in adjacent table entries
    std::unordered_multimap<uint64_t, uint32_t> kpair_idx;
    auto kpair_field = [&](int32_t v1, int32_t v2) -> uint32_t {
        const uint64_t b1 = g.nodes[v1].bits, b2 = g.nodes[v2].bits;
        const uint64_t key = b1 * 0x9E3779B97F4A7C15ull ^ (b2 + 0x7F4A7C159E3779B9ull + (b1 << 6) + (b1 >> 2));
        for (auto range = kpair_idx.equal_range(key); range.first != range.second; ++range.first) {
            uint32_t i = range.first->second;
            uint64_t c1, c2;
            std::memcpy(&c1, &out.consts[i], 8);
            std::memcpy(&c2, &out.consts[i + 1], 8);
            if (c1 == b1 && c2 == b2) return make_field(KIND_K, i & (MAX_INDEX - 1));
        }
        const uint32_t i = (uint32_t)out.consts.size();
        double d1, d2;
        std::memcpy(&d1, &b1, 8);
        std::memcpy(&d2, &b2, 8);
        out.consts.push_back(d1);
        out.consts.push_back(d2);
        kpair_idx.emplace(key, i);
        return make_field(KIND_K, i & (MAX_INDEX - 1));
    };

    out.code.clear();
    out.code.reserve(2 * (size_t)NE + 2);
    bool overflow = false;
    auto sfield = [&](int32_t sl) { return make_field(KIND_S, (uint32_t)sl & (MAX_INDEX - 1)); };
    for (int32_t j = 0; j < NE; ++j) {
        const Emit &e = E[j];
        uint32_t f[3] = {FIELD_NONE, FIELD_NONE, FIELD_NONE};
        for (int k = 0; k < 3; ++k) {
            int32_t s = e.src[k];
            if (s == SRC_LOADK) { f[k] = FIELD_ACC; continue; }
            if (s < 0) continue;
            if (g.nodes[s].op == V_CONST) f[k] = kfield(s);
            else if (j > 0 && E[j - 1].node == s && is_instr_value(s)) f[k] = FIELD_ACC;
            else f[k] = sfield(slot[s]);
        }
        // sources dying here release their slots before the destination is placed: the kernel
        // reads every source before it writes (same contract as macros.rs, fusion.rs:742-753)
        for (int32_t v : dying[j]) free_slots.push_back(slot[v]);
        uint32_t fd = FIELD_ACC;
        if (e.node >= 0 && needs_slot[e.node]) {
            slot[e.node] = alloc();
            fd = sfield(slot[e.node]);
        } else if (e.node < 0) {
            fd = FIELD_NONE;
        }
        const Node *n = e.node >= 0 ? &g.nodes[e.node] : nullptr;
        uint32_t fa = f[0], fb = f[1], fc = f[2], imm = 0;
        const uint32_t op = e.op;
        uint32_t xop;
        if (op == D_LOADK) {
            xop = X_LOADK;
        } else if (op >= D_FIRST_LIGHT_UNARY && op < D_FIRST_LIGHT_UNARY + D_N_LIGHT_UNARY) {
            xop = X_ULIGHT + 2 * (op - D_FIRST_LIGHT_UNARY) + (field_kind(fa) == KIND_ACC ? 1 : 0);
        } else if (op >= D_FIRST_HEAVY_UNARY && op < D_FIRST_HEAVY_UNARY + D_N_HEAVY_UNARY) {
            xop = X_UHEAVY + (op - D_FIRST_HEAVY_UNARY);
            if (op == D_SINCOS) {
                fc = FIELD_NONE;
                if (e.cosval >= 0 && needs_slot[e.cosval]) {
                    slot[e.cosval] = alloc();
                    fc = sfield(slot[e.cosval]);
                }
            } else if (op == D_BUILTIN1 || op == D_BUILTIN3) {
                imm = n->imm;
            }
        } else if (op >= D_FIRST_LIGHT_BINARY && op < D_FIRST_LIGHT_BINARY + D_N_LIGHT_BINARY) {
            xop = X_BLIGHT + 9 * (op - D_FIRST_LIGHT_BINARY) + 3 * field_kind(fa) + field_kind(fb);
        } else if (op >= D_FIRST_HEAVY_BINARY && op < D_FIRST_HEAVY_BINARY + D_N_HEAVY_BINARY) {
            xop = X_BHEAVY + (op - D_FIRST_HEAVY_BINARY);
            if (op == D_POWI) {
                double nd;
                std::memcpy(&nd, &g.nodes[n->b].bits, 8);
                imm = (uint32_t)(int32_t)nd;
            } else if (op == D_BUILTIN2 || op == D_BUILTIN4) {
                imm = n->imm;
            }
        } else {
            xop = X_TERN + (op - D_FIRST_TERNARY);
        }
        if (e.prefix == PREFIX_MUL) fb = kfield(e.pk1);
        else if (e.prefix == PREFIX_FMA) fb = kpair_field(e.pk1, e.pk2);
        if (e.prefix != PREFIX_NONE) ++out.n_prefix_fusions;
        out.code.push_back(make_lo(xop, fd, fa, fb, e.prefix));
        if (xop_has_hi(xop)) out.code.push_back(make_hi(fc, imm));
        ++out.n_dev_instructions;
        if (n_slots > MAX_INDEX || out.consts.size() > MAX_INDEX) overflow = true;
    }
    out.code.push_back(make_lo(X_END, FIELD_NONE, FIELD_NONE, FIELD_NONE));

    out.result_field.clear();
    for (int32_t r : results) {
        if (g.nodes[r].op == V_CONST) out.result_field.push_back(kfield(r));
        else out.result_field.push_back(make_field(KIND_S, (uint32_t)slot[r] & (MAX_INDEX - 1)));
    }
    if (n_slots > MAX_INDEX || out.consts.size() > MAX_INDEX || overflow) {
        std::ostringstream os;
        os << "program needs " << n_slots << " live-value slots and " << out.consts.size()
           << " constants; the device encoding holds " << MAX_INDEX << " of each";
        err = os.str();
        return SAF_ERR_UNSUPPORTED;
    }
    out.n_slots = n_slots ? n_slots : 1;
    out.quirk_flags = g.quirks;
    return SAF_OK;
    }; // emit_program

    // ---- schedule: two candidate orders, keep the one that is cheaper on chip ----
    //  (a) progress-normalised lockstep of all long sums   (b) demand-driven + eager
    // Fewer live-value slots means more resident threads; more accumulator forwards means less
    // shared-memory traffic.  Slots win when they differ by more than a quarter.
    DeviceProgram cand_a, cand_b;
    std::string err_a, err_b;
    int rc_a, rc_b;
    {
        std::string saved = err;
        const std::vector<int32_t> base_a = schedule_by_progress(g, live, results);
        rc_a = emit_program(base_a, cand_a);
        err_a = err;
        err = saved;
        // chain following keeps (a)'s lockstep of the long sums and adds accumulator forwards inside a term
        DeviceProgram cand_a2;
        const int rc_a2 = emit_program(follow_chains(g, live, results, base_a), cand_a2);
        if (rc_a2 == SAF_OK && (rc_a != SAF_OK || (cand_a2.n_slots <= cand_a.n_slots + cand_a.n_slots / 8 &&
                                                   cand_a2.n_acc_forwards > cand_a.n_acc_forwards))) {
            cand_a = std::move(cand_a2);
            rc_a = rc_a2;
            err_a = err;
        }
        err = saved;
        rc_b = emit_program(schedule_nodes(g, live, results), cand_b);
        err_b = err;
        err = saved;
        // (c) term-aligned merge of the long sums (+ chain following); replaces (a) when it needs clearly fewer slots
        // (more resident warps: the interpreter is latency-bound at the occupancy big register files leave)
        const char *no_merge = getenv("SAF_SCHED");
        if (!(no_merge && no_merge[0] == 'n')) {
            DeviceProgram cand_c, cand_c2;
            const std::vector<int32_t> base_c = schedule_spine_merge(g, live, results);
            int rc_c = emit_program(base_c, cand_c);
            const int rc_c2 = emit_program(follow_chains(g, live, results, base_c), cand_c2);
            if (rc_c2 == SAF_OK && (rc_c != SAF_OK || (cand_c2.n_slots <= cand_c.n_slots &&
                                                       cand_c2.n_acc_forwards > cand_c.n_acc_forwards))) {
                cand_c = std::move(cand_c2);
                rc_c = rc_c2;
            }
            const bool force_c = no_merge && no_merge[0] == 'm';
            if (rc_c == SAF_OK && (force_c || rc_a != SAF_OK || cand_c.n_slots * 8 <= cand_a.n_slots * 7)) {
                cand_a = std::move(cand_c);
                rc_a = rc_c;
                err_a = err;
            }
            err = saved;
        }
    }
    const char *force = getenv("SAF_SCHED");
    bool pick_a;
    if (force && force[0] == 'p') pick_a = true;
    else if (force && force[0] == 'd') pick_a = false;
    else if (rc_a != SAF_OK) pick_a = false;
    else if (rc_b != SAF_OK) pick_a = true;
    else if (cand_a.n_slots * 4 < cand_b.n_slots * 3) pick_a = true;
    else if (cand_b.n_slots * 4 < cand_a.n_slots * 3) pick_a = false;
    else pick_a = cand_a.n_acc_forwards > cand_b.n_acc_forwards; // ties: the demand-driven order measured faster
    if (pick_a) {
        out = std::move(cand_a);
        err = err_a;
        return rc_a;
    }
    out = std::move(cand_b);
    err = err_b;
    return rc_b;
}

// ------------------------------------------------------------------------------------------------
static const char *DEV_NAMES[D_OPCOUNT] = {
    "end", "loadk", "copy", "neg", "square", "cube", "pow4", "pow3_2", "invpow3_2", "invsqrt", "invsquare",
    "invcube", "recip", "sqrt", "sin", "cos", "exp", "ln", "recipexpm1", "expsqr", "expsqrneg", "expneg",
    "sincos", "builtin1", "builtin3", "add", "sub", "mul", "negmul", "div", "pow", "powi", "builtin2",
    "builtin4", "arg2", "fma", "fms", "fnma", "fnms"};

static void fmt_field(std::ostringstream &os, uint32_t f) {
    uint32_t kind = field_kind(f), idx = field_index(f);
    if (kind == KIND_S) os << "S" << idx;
    else if (kind == KIND_K) os << "K" << idx;
    else if (kind == KIND_ACC) os << "ACC";
    else os << "-";
}

std::string disassemble(const DeviceProgram &p) {
    std::ostringstream os;
    const size_t n_instr = p.n_dev_instructions;
    os << "=== Device program ===\n";
    os << "params " << p.param_count << ", results " << p.result_field.size() << ", slots " << p.n_slots
       << ", constants " << p.consts.size() << ", instructions " << n_instr
       << ", acc-forwards " << p.n_acc_forwards << ", fp64 slots/point " << p.fp64_slots << "\n";
    for (size_t j = 0; j < p.param_slot.size(); ++j) {
        os << "  param " << j << " -> ";
        if (p.param_slot[j] == NO_SLOT) os << "(unused)\n";
        else os << "S" << p.param_slot[j] << "\n";
    }
    for (size_t i = 0; i < p.consts.size(); ++i) {
        os.precision(17);
        os << "  K" << i << " = " << p.consts[i] << "\n";
    }
    for (size_t i = 0, pc = 0; i < n_instr; ++i) {
        const uint64_t lo = p.code[pc++];
        const uint64_t hi = xop_has_hi((uint32_t)(lo & 0xff)) ? p.code[pc++] : make_hi(FIELD_NONE, 0);
        const uint32_t x = lo & 0xff, d = (lo >> 16) & 0xffff, a = (lo >> 32) & 0xffff, b = (lo >> 48) & 0xffff,
                       c = hi & 0xffff, imm = (uint32_t)(hi >> 32);
        uint32_t op, ka, kb;
        decode_xop(x, &op, &ka, &kb);
        os << "  [" << i << "] ";
        fmt_field(os, d);
        const uint32_t prefix = (lo >> 8) & 0xff;
        os << " = " << (op < D_OPCOUNT ? DEV_NAMES[op] : "?") << "(";
        fmt_field(os, a);
        if (prefix == PREFIX_MUL) {
            os << " * ";
            fmt_field(os, b);
        } else if (prefix == PREFIX_FMA) {
            os << " * ";
            fmt_field(os, b);
            os << " + K" << field_index(b) + 1;
        } else if (field_kind(b) != KIND_NONE) {
            os << ", ";
            fmt_field(os, b);
        }
        if (field_kind(c) != KIND_NONE) {
            os << (op == D_SINCOS ? "; cos -> " : ", ");
            fmt_field(os, c);
        }
        if (op == D_POWI) os << ", n=" << (int32_t)imm;
        if (op == D_BUILTIN1 || op == D_BUILTIN2 || op == D_BUILTIN3 || op == D_BUILTIN4) os << ", fn=" << imm;
        os << ")\n";
    }
    for (size_t k = 0; k < p.result_field.size(); ++k) {
        os << "  result " << k << " <- ";
        fmt_field(os, p.result_field[k]);
        os << "\n";
    }
    return os.str();
}

} // namespace saf
