// compile.cpp -- expression string -> SymbAnaFis bytecode, in C++.
//
// The reference compiles with a 7 kLoC Rust pipeline (src/evaluator/logic/bytecode/compile/**); without a Rust toolchain
// programs used to come from the Python harness (symbanafis_b200/compiler.py + expr.py).  This file is the same
// compiler behind the C ABI (saf_compile_exprs), so that a C / C++ / Rust host gets a device program from expression
// strings without Python: hash-consed expression DAG with the constant-folding constructors of the parser
// (src/parser/logic/pratt.rs:87-120: `a - b` is Sum[a, Product[-1, b]]), Pratt parser, and the documented lowering rules
//   sums      -> MulAdd / MulSub / NegMulAdd / Sub / Add..AddN        codegen/lower/sum.rs:104-344
//   products  -> constants folded into one trailing factor, Mul..MulN  codegen/lower/product.rs:10-91
//   division  -> RecipExpm1, u/c -> u*(1/c), Sinc, Recip, Div          codegen/lower/div.rs:13-68
//   powers    -> Square .. InvPow3_2, Powi, Pow                        codegen/lower/pow.rs:186-334
//   functions -> ExpSqr / ExpSqrNeg / Builtin1{ExpNeg}, FnOp table     codegen/lower/func.rs:14-110, vir/registry.rs:14-160
//   Sin + Cos of one argument -> SinCos                                optimize/fusion.rs:105-187
//   value numbering while emitting, backward dead-code elimination, constant compaction (optimize/compact.rs:10-62),
//   linear-scan register allocation with a LIFO free list              emit/reg_alloc.rs:106-138,441-449
//   flat encoding                                                      emit/assemble.rs:9-103
// It produces, bit for bit, what compiler.py produces (tests/test_cpp_compiler.py compares the two over the five
// workloads, the reference's SIMD-surface corpus and random expressions); like compiler.py it follows the reference's
// rules, not its exact register numbering (no Rust build to compare with: "parity unpinned" on the byte level).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/saf_b200.h"
#include "lower.h"
#include "saf_isa.h"

namespace saf {
namespace cc {

namespace {

constexpr double EPSILON = 1e-14; // crate::EPSILON, src/lib.rs:321

// ---- expression DAG (symbanafis_b200/expr.py) -------------------------------------------------------------------
enum Kind : uint8_t { NUM, SYM, CALL, SUM, PROD, DIV, POW };
struct Node {
    Kind kind;
    double value = 0.0;
    std::string name;
    std::vector<int> args;
};

struct Dag {
    std::vector<Node> nodes;
    std::map<std::string, int> intern;

    int make(Kind kind, double value, const std::string &name, const std::vector<int> &args) {
        std::string key;
        if (kind == NUM) {
            uint64_t bits;
            if (value != value) bits = 0x7ff8000000000000ull; // every NaN is one node
            else std::memcpy(&bits, &value, 8);
            key = "n" + std::to_string(bits);
        } else {
            key = std::string(1, (char)('a' + kind)) + name + "(";
            for (int a : args) key += std::to_string(a) + ",";
        }
        auto it = intern.find(key);
        if (it != intern.end()) return it->second;
        Node n;
        n.kind = kind;
        n.value = value;
        n.name = name;
        n.args = args;
        nodes.push_back(n);
        return intern[key] = (int)nodes.size() - 1;
    }
    int number(double v) { return make(NUM, v, "", {}); }
    int symbol(const std::string &s) { return make(SYM, 0.0, s, {}); }
    int call(const std::string &s, const std::vector<int> &a) { return make(CALL, 0.0, s, a); }
    int sum(const std::vector<int> &terms) {
        std::vector<int> flat;
        double c = 0.0;
        bool has_const = false;
        for (int t : terms) {
            std::vector<int> inner = nodes[t].kind == SUM ? nodes[t].args : std::vector<int>{t};
            for (int u : inner) {
                if (nodes[u].kind == NUM) {
                    c += nodes[u].value;
                    has_const = true;
                } else
                    flat.push_back(u);
            }
        }
        if (has_const && (c != 0.0 || flat.empty())) flat.push_back(number(c));
        if (flat.empty()) return number(0.0);
        if (flat.size() == 1) return flat[0];
        return make(SUM, 0.0, "", flat);
    }
    int product(const std::vector<int> &factors) {
        std::vector<int> flat;
        double coeff = 1.0;
        for (int f : factors) {
            std::vector<int> inner = nodes[f].kind == PROD ? nodes[f].args : std::vector<int>{f};
            for (int u : inner) {
                if (nodes[u].kind == NUM) coeff *= nodes[u].value;
                else flat.push_back(u);
            }
        }
        if (coeff == 0.0) return number(0.0);
        if (flat.empty()) return number(coeff);
        if (coeff != 1.0) flat.insert(flat.begin(), number(coeff));
        if (flat.size() == 1) return flat[0];
        return make(PROD, 0.0, "", flat);
    }
    int div(int a, int b) {
        const Node &na = nodes[a], &nb = nodes[b];
        if (nb.kind == NUM && nb.value == 1.0) return a;
        if (na.kind == NUM && nb.kind == NUM && nb.value != 0.0) return number(na.value / nb.value);
        if (na.kind == NUM && na.value == 0.0) return a;
        return make(DIV, 0.0, "", {a, b});
    }
    int pow(int a, int b) {
        if (nodes[b].kind == NUM) {
            const double e = nodes[b].value;
            if (e == 1.0) return a;
            if (e == 0.0) return number(1.0);
            if (nodes[a].kind == NUM) {
                const double v = std::pow(nodes[a].value, e);
                if (std::isfinite(v)) return number(v);
            }
        }
        return make(POW, 0.0, "", {a, b});
    }
    int sub(int a, int b) { return sum({a, product({number(-1.0), b})}); }
    int neg(int a) { return product({number(-1.0), a}); }

    // iterative post-order, each node once, arguments left to right (expr.py: postorder)
    std::vector<int> postorder(int root) const {
        std::vector<int> out;
        std::vector<uint8_t> seen(nodes.size(), 0);
        std::vector<std::pair<int, bool>> stack{{root, false}};
        while (!stack.empty()) {
            auto [n, done] = stack.back();
            stack.pop_back();
            if (done) {
                out.push_back(n);
                continue;
            }
            if (seen[n]) continue;
            seen[n] = 1;
            stack.push_back({n, true});
            const auto &a = nodes[n].args;
            for (size_t i = a.size(); i-- > 0;)
                if (!seen[a[i]]) stack.push_back({a[i], false});
        }
        return out;
    }
};

// ---- Pratt parser (expr.py: _Parser; binding powers + - 10, * / 20, unary - 25, ^ 30 right associative) ----------
struct Tok {
    char kind; // 'n' number, 'i' identifier, 'o' operator, 'e' end
    std::string val;
};

bool tokenize(const std::string &src_in, std::vector<Tok> &toks, std::string &err) {
    std::string src = src_in;
    while (!src.empty() && std::isspace((unsigned char)src.back())) src.pop_back();
    size_t pos = 0;
    auto digit = [&](size_t p) { return p < src.size() && std::isdigit((unsigned char)src[p]); };
    while (pos < src.size()) {
        while (pos < src.size() && std::isspace((unsigned char)src[pos])) ++pos;
        if (pos >= src.size()) break;
        const char ch = src[pos];
        size_t q = pos;
        if (digit(pos) || (ch == '.' && digit(pos + 1))) {
            if (digit(q)) {
                while (digit(q)) ++q;
                if (q < src.size() && src[q] == '.') ++q;
                while (digit(q)) ++q;
            } else {
                ++q;
                while (digit(q)) ++q;
            }
            if (q < src.size() && (src[q] == 'e' || src[q] == 'E')) { // exponent only when digits follow
                size_t r = q + 1;
                if (r < src.size() && (src[r] == '+' || src[r] == '-')) ++r;
                if (digit(r)) {
                    while (digit(r)) ++r;
                    q = r;
                }
            }
            toks.push_back({'n', src.substr(pos, q - pos)});
        } else if (std::isalpha((unsigned char)ch) || ch == '_') {
            while (q < src.size() && (std::isalnum((unsigned char)src[q]) || src[q] == '_')) ++q;
            toks.push_back({'i', src.substr(pos, q - pos)});
        } else if (ch == '*' && pos + 1 < src.size() && src[pos + 1] == '*') {
            toks.push_back({'o', "^"});
            q = pos + 2;
        } else if (std::strchr("-+*/^(),", ch)) {
            toks.push_back({'o', std::string(1, ch)});
            q = pos + 1;
        } else {
            err = std::string("unexpected character '") + ch + "' at " + std::to_string(pos);
            return false;
        }
        pos = q;
    }
    toks.push_back({'e', ""});
    return true;
}

struct Parser {
    Dag &g;
    std::vector<Tok> toks;
    size_t i = 0;
    std::string err;
    explicit Parser(Dag &dag) : g(dag) {}
    const Tok &peek() const { return toks[i]; }
    Tok next() { return toks[i++]; }
    bool is_op(const Tok &t, const char *v) const { return t.kind == 'o' && t.val == v; }
    int fail(const std::string &m) {
        if (err.empty()) err = m;
        return -1;
    }
    int expr(int min_bp) {
        Tok t = next();
        int lhs;
        if (t.kind == 'n') {
            lhs = g.number(std::strtod(t.val.c_str(), nullptr));
        } else if (t.kind == 'i') {
            if (is_op(peek(), "(")) {
                next();
                std::vector<int> args;
                if (!is_op(peek(), ")")) {
                    for (;;) {
                        const int a = expr(0);
                        if (a < 0) return -1;
                        args.push_back(a);
                        if (is_op(peek(), ",")) {
                            next();
                            continue;
                        }
                        break;
                    }
                }
                if (!is_op(peek(), ")")) return fail("expected ')', got '" + peek().val + "'");
                next();
                lhs = g.call(t.val, args);
            } else
                lhs = g.symbol(t.val);
        } else if (is_op(t, "(")) {
            lhs = expr(0);
            if (lhs < 0) return -1;
            if (!is_op(peek(), ")")) return fail("expected ')', got '" + peek().val + "'");
            next();
        } else if (is_op(t, "-")) {
            const int rhs = expr(25);
            if (rhs < 0) return -1;
            lhs = g.neg(rhs);
        } else if (is_op(t, "+")) {
            lhs = expr(25);
            if (lhs < 0) return -1;
        } else
            return fail("unexpected token '" + t.val + "'");

        for (;;) {
            const Tok &p = peek();
            if (p.kind == 'e' || is_op(p, ")") || is_op(p, ",")) break;
            if (p.kind == 'o') {
                const std::string v = p.val;
                if (v == "+" || v == "-") {
                    if (10 < min_bp) break;
                    next();
                    const int rhs = expr(11);
                    if (rhs < 0) return -1;
                    lhs = v == "+" ? g.sum({lhs, rhs}) : g.sub(lhs, rhs);
                } else if (v == "*" || v == "/") {
                    if (20 < min_bp) break;
                    next();
                    const int rhs = expr(21);
                    if (rhs < 0) return -1;
                    lhs = v == "*" ? g.product({lhs, rhs}) : g.div(lhs, rhs);
                } else if (v == "^") {
                    if (30 < min_bp) break;
                    next();
                    const int rhs = expr(30); // right associative
                    if (rhs < 0) return -1;
                    lhs = g.pow(lhs, rhs);
                } else if (v == "(") { // implicit multiplication: 3(x+1)
                    if (20 < min_bp) break;
                    const int rhs = expr(21);
                    if (rhs < 0) return -1;
                    lhs = g.product({lhs, rhs});
                } else
                    return fail("unexpected operator '" + v + "'");
            } else { // implicit multiplication: 2x, 2 sin(x)
                if (20 < min_bp) break;
                const int rhs = expr(21);
                if (rhs < 0) return -1;
                lhs = g.product({lhs, rhs});
            }
        }
        return lhs;
    }
    int parse(const std::string &src) {
        if (!tokenize(src, toks, err)) return -1;
        const int e = expr(0);
        if (e < 0) return -1;
        if (peek().kind != 'e') return fail("unexpected token '" + peek().val + "'");
        return e;
    }
};

// ---- FnOp table (bytecode/functions.rs:38-117) --------------------------------------------------------------------
struct FnEntry {
    const char *name;
    int arity;
};
const FnEntry FNOPS[] = {
    {"sin", 1}, {"cos", 1}, {"tan", 1}, {"cot", 1}, {"sec", 1}, {"csc", 1}, {"asin", 1}, {"acos", 1}, {"atan", 1}, {"acot", 1},
    {"asec", 1}, {"acsc", 1}, {"sinh", 1}, {"cosh", 1}, {"tanh", 1}, {"coth", 1}, {"sech", 1}, {"csch", 1}, {"asinh", 1},
    {"acosh", 1}, {"atanh", 1}, {"acoth", 1}, {"acsch", 1}, {"asech", 1}, {"exp", 1}, {"expm1", 1}, {"exp_neg", 1}, {"ln", 1},
    {"log1p", 1}, {"sqrt", 1}, {"cbrt", 1}, {"abs", 1}, {"signum", 1}, {"floor", 1}, {"ceil", 1}, {"round", 1}, {"erf", 1},
    {"erfc", 1}, {"gamma", 1}, {"lgamma", 1}, {"digamma", 1}, {"trigamma", 1}, {"tetragamma", 1}, {"sinc", 1},
    {"lambert_w", 1}, {"elliptic_k", 1}, {"elliptic_e", 1}, {"zeta", 1}, {"exp_polar", 1}, {"atan2", 2}, {"log", 2},
    {"bessel_j", 2}, {"bessel_y", 2}, {"bessel_i", 2}, {"bessel_k", 2}, {"polygamma", 2}, {"beta", 2}, {"zeta_deriv", 2},
    {"hermite", 2}, {"assoc_legendre", 3}, {"spherical_harmonic", 4}};
int fnop_of(const std::string &name) {
    for (size_t i = 0; i < sizeof FNOPS / sizeof FNOPS[0]; ++i)
        if (name == FNOPS[i].name) return (int)i;
    return -1;
}
std::string fn_alias(const std::string &n) { // vir/registry.rs:14-87
    static const std::map<std::string, std::string> A = {
        {"sign", "signum"}, {"sgn", "signum"}, {"ynm", "spherical_harmonic"}, {"lambertw", "lambert_w"},
        {"besselj", "bessel_j"}, {"bessely", "bessel_y"}, {"besseli", "bessel_i"}, {"besselk", "bessel_k"},
        {"ellipk", "elliptic_k"}, {"ellipe", "elliptic_e"}, {"loggamma", "lgamma"}};
    auto it = A.find(n);
    return it == A.end() ? n : it->second;
}
bool const_fold(const std::string &name, double c, double *out) { // vir/registry.rs:89-160
    double v;
    if (name == "sin") v = std::sin(c);
    else if (name == "cos") v = std::cos(c);
    else if (name == "tan") v = std::tan(c);
    else if (name == "exp") v = std::exp(c);
    else if (name == "sqrt") v = std::sqrt(c);
    else if (name == "ln") v = std::log(c);
    else if (name == "abs") v = std::fabs(c);
    else if (name == "sinh") v = std::sinh(c);
    else if (name == "cosh") v = std::cosh(c);
    else if (name == "tanh") v = std::tanh(c);
    else if (name == "asin") v = std::asin(c);
    else if (name == "acos") v = std::acos(c);
    else if (name == "atan") v = std::atan(c);
    else return false;
    if (!std::isfinite(v)) return false;
    *out = v;
    return true;
}
bool known_constant(const std::string &n, double *v) {
    if (n == "pi" || n == "PI" || n == "Pi") { *v = 3.141592653589793; return true; }
    if (n == "e" || n == "E") { *v = 2.718281828459045; return true; }
    return false;
}

// ---- virtual instructions -----------------------------------------------------------------------------------------
struct VReg {
    char kind = 't'; // 'c' constant, 'p' parameter, 't' temporary
    uint32_t idx = 0;
    bool operator==(const VReg &o) const { return kind == o.kind && idx == o.idx; }
    bool operator<(const VReg &o) const { return kind != o.kind ? kind < o.kind : idx < o.idx; }
};
enum VOp : uint8_t { // n-ary Add / Mul; the rest maps to one reference opcode
    VADD, VMUL, VNEG, VNEGMUL, VMULADD, VMULSUB, VNEGMULADD, VSUB, VDIV, VRECIP, VSQUARE, VCUBE, VPOW4, VINVSQUARE, VINVCUBE,
    VPOWI, VSQRT, VINVSQRT, VPOW3_2, VINVPOW3_2, VPOW, VEXP, VEXPSQR, VEXPSQRNEG, VSIN, VCOS, VLN, VRECIPEXPM1, VSINCOS,
    VBUILTIN
};
struct VInstr {
    VOp op;
    std::vector<VReg> dests, srcs;
    int64_t extra = 0;
    bool has_extra = false;
};

struct Gen {
    const Dag &g;
    std::vector<std::string> params;
    std::map<std::string, uint32_t> param_idx;
    std::vector<double> consts;
    std::map<uint64_t, uint32_t> const_idx;
    std::vector<VInstr> instrs;
    uint32_t n_temps = 0;
    std::map<std::string, std::vector<VReg>> gvn;
    std::string err;
    std::vector<VReg> m; // node -> vreg
    std::vector<uint8_t> have;

    Gen(const Dag &dag, const std::vector<std::string> &ps) : g(dag), params(ps) {
        for (size_t i = ps.size(); i-- > 0;) param_idx[ps[i]] = (uint32_t)i; // the first of two equal names wins
        m.resize(dag.nodes.size());
        have.assign(dag.nodes.size(), 0);
    }
    VReg add_const(double v) {
        uint64_t bits;
        std::memcpy(&bits, &v, 8);
        auto it = const_idx.find(bits);
        uint32_t i;
        if (it == const_idx.end()) {
            i = (uint32_t)consts.size();
            consts.push_back(v);
            const_idx[bits] = i;
        } else
            i = it->second;
        return VReg{'c', i};
    }
    bool cv(int node, double *v) const {
        const VReg &r = m[node];
        if (r.kind != 'c') return false;
        *v = consts[r.idx];
        return true;
    }
    bool cvr(const VReg &r, double *v) const {
        if (r.kind != 'c') return false;
        *v = consts[r.idx];
        return true;
    }
    std::vector<VReg> emit_n(VOp op, const std::vector<VReg> &srcs, bool has_extra, int64_t extra, int n_dest) {
        std::vector<VReg> ks = srcs;
        if ((op == VADD || op == VMUL) && ks.size() == 2 && ks[1] < ks[0]) std::swap(ks[0], ks[1]); // commutative key
        std::string key = std::to_string((int)op) + "|";
        for (const VReg &r : ks) key += std::string(1, r.kind) + std::to_string(r.idx) + ",";
        key += has_extra ? "|" + std::to_string(extra) : "|-";
        auto it = gvn.find(key);
        if (it != gvn.end()) return it->second;
        VInstr in;
        in.op = op;
        in.srcs = srcs;
        in.extra = extra;
        in.has_extra = has_extra;
        for (int k = 0; k < n_dest; ++k) in.dests.push_back(VReg{'t', n_temps + (uint32_t)k});
        n_temps += (uint32_t)n_dest;
        instrs.push_back(in);
        return gvn[key] = in.dests;
    }
    VReg emit(VOp op, const std::vector<VReg> &srcs) { return emit_n(op, srcs, false, 0, 1)[0]; }
    VReg emit_x(VOp op, const std::vector<VReg> &srcs, int64_t extra) { return emit_n(op, srcs, true, extra, 1)[0]; }
    VReg add_vregs(const std::vector<VReg> &v) {
        if (v.empty()) return add_const(0.0);
        if (v.size() == 1) return v[0];
        return emit(VADD, v);
    }
    VReg mul_vregs(const std::vector<VReg> &v) {
        if (v.empty()) return add_const(1.0);
        if (v.size() == 1) return v[0];
        if (v.size() == 2 && v[0] == v[1]) return emit(VSQUARE, {v[0]});
        return emit(VMUL, v);
    }

    // products ------------------------------------------------------------------------------------------------------
    void split_product(int e, double *c, bool *has, std::vector<VReg> *v) const {
        *c = 1.0;
        *has = false;
        v->clear();
        for (int f : g.nodes[e].args) {
            double fc;
            if (cv(f, &fc)) {
                *c *= fc;
                *has = true;
            } else
                v->push_back(m[f]);
        }
    }
    VReg lower_product(int n) {
        double c;
        bool has;
        std::vector<VReg> v;
        split_product(n, &c, &has, &v);
        if (std::isfinite(c)) {
            if (c == 0.0) c = 0.0; // -0.0 -> +0.0 (product.rs:67-71)
            if (std::fabs(c - 1.0) > EPSILON) {
                if (v.empty()) return add_const(c);
                if (std::fabs(c + 1.0) < EPSILON && v.size() == 1) return emit(VNEG, v);
                if (std::fabs(c + 1.0) < EPSILON && v.size() == 2) return emit(VNEGMUL, v);
                v.push_back(add_const(c)); // the constant goes last (product.rs:62-73)
            }
        } else
            v.push_back(add_const(c));
        return mul_vregs(v);
    }
    // try_extract_negated_product (sum.rs:36-68): vreg of e without its -1 factor
    bool neg_inner(int e, VReg *out) {
        if (g.nodes[e].kind != PROD) return false;
        double c;
        bool has;
        std::vector<VReg> v;
        split_product(e, &c, &has, &v);
        if (!has || !std::isfinite(c) || c >= 0.0) return false;
        const double pos = -c;
        if (std::fabs(pos - 1.0) > EPSILON) v.push_back(add_const(pos));
        *out = mul_vregs(v);
        return true;
    }
    bool prod2(int e, VReg *a, VReg *b) {
        if (g.nodes[e].kind != PROD) return false;
        double c;
        bool has;
        std::vector<VReg> v;
        split_product(e, &c, &has, &v);
        if (!std::isfinite(c)) return false;
        if (!has || std::fabs(c - 1.0) <= EPSILON) {
            if (v.size() != 2) return false;
            *a = v[0];
            *b = v[1];
            return true;
        }
        if (c > 0 && v.size() == 1) {
            *a = v[0];
            *b = add_const(c);
            return true;
        }
        return false;
    }
    bool negprod2(int e, VReg *a, VReg *b) {
        if (g.nodes[e].kind != PROD) return false;
        double c;
        bool has;
        std::vector<VReg> v;
        split_product(e, &c, &has, &v);
        if (!has || !std::isfinite(c) || c >= 0.0) return false;
        if (std::fabs(c + 1.0) <= EPSILON) {
            if (v.size() != 2) return false;
            *a = v[0];
            *b = v[1];
            return true;
        }
        if (v.size() == 1) {
            *a = v[0];
            *b = add_const(-c);
            return true;
        }
        return false;
    }

    // sums ----------------------------------------------------------------------------------------------------------
    VReg lower_sum(int n) {
        const std::vector<int> &terms = g.nodes[n].args;
        VReg a, b, ng;
        if (terms.size() == 2) {
            const int t0 = terms[0], t1 = terms[1];
            double c0 = 0, c1 = 0;
            const bool h0 = cv(t0, &c0), h1 = cv(t1, &c1);
            if (h0 && h1 && std::isfinite(c0 + c1)) return add_const(c0 + c1);
            if (h0 && std::isfinite(c0) && std::fabs(c0) < EPSILON) return m[t1];
            if (h1 && std::isfinite(c1) && std::fabs(c1) < EPSILON) return m[t0];
            const int order[2][2] = {{t0, t1}, {t1, t0}};
            for (auto &o : order) // a*b - c -> MulSub
                if (!h0 && !h1 && prod2(o[0], &a, &b) && neg_inner(o[1], &ng)) return emit(VMULSUB, {a, b, ng});
            for (auto &o : order) // c - a*b -> NegMulAdd
                if (negprod2(o[0], &a, &b)) return emit(VNEGMULADD, {a, b, m[o[1]]});
            for (auto &o : order) // a*b + c -> MulAdd
                if (prod2(o[0], &a, &b)) return emit(VMULADD, {a, b, m[o[1]]});
            for (auto &o : order) // x - y -> Sub
                if (neg_inner(o[0], &ng)) return emit(VSUB, {m[o[1]], ng});
            return emit(VADD, {m[t0], m[t1]});
        }
        std::vector<VReg> pos, neg;
        double cacc = 0.0;
        bool has_c = false, has_pp = false, has_np = false;
        VReg pp0, pp1, np0, np1;
        for (int t : terms) {
            double c;
            if (cv(t, &c)) {
                cacc += c;
                has_c = true;
                continue;
            }
            if (negprod2(t, &a, &b)) {
                if (!has_np) { np0 = a; np1 = b; has_np = true; }
                else neg.push_back(mul_vregs({a, b}));
                continue;
            }
            if (prod2(t, &a, &b)) {
                if (!has_pp) { pp0 = a; pp1 = b; has_pp = true; }
                else pos.push_back(mul_vregs({a, b}));
                continue;
            }
            if (neg_inner(t, &ng)) neg.push_back(ng);
            else pos.push_back(m[t]);
        }
        if (has_c) {
            if (std::isfinite(cacc)) {
                if (std::fabs(cacc) > EPSILON) (cacc > 0 ? pos : neg).push_back(add_const(std::fabs(cacc)));
            } else
                pos.push_back(add_const(cacc));
        }
        // assemble_nary_sum (sum.rs:261-344)
        if (has_pp && !has_np && neg.empty()) {
            if (pos.empty()) return mul_vregs({pp0, pp1});
            return emit(VMULADD, {pp0, pp1, add_vregs(pos)});
        }
        if (!has_pp && !has_np && neg.empty()) return add_vregs(pos);
        if (!has_pp && has_np && pos.empty()) {
            VReg inner;
            if (neg.empty()) inner = mul_vregs({np0, np1});
            else {
                std::vector<VReg> nn = neg;
                nn.push_back(mul_vregs({np0, np1}));
                inner = add_vregs(nn);
            }
            return emit(VNEG, {inner});
        }
        if (!has_pp && !has_np && pos.empty()) return emit(VNEG, {add_vregs(neg)});
        if (has_pp && !has_np && pos.empty()) return emit(VMULSUB, {pp0, pp1, add_vregs(neg)});
        if (!has_pp && has_np && neg.empty()) return emit(VNEGMULADD, {np0, np1, add_vregs(pos)});
        bool has_pos_v = true;
        VReg pos_v;
        if (has_pp) pos_v = pos.empty() ? mul_vregs({pp0, pp1}) : emit(VMULADD, {pp0, pp1, add_vregs(pos)});
        else if (!pos.empty()) pos_v = add_vregs(pos);
        else has_pos_v = false;
        if (has_np) neg.push_back(mul_vregs({np0, np1}));
        const VReg neg_v = add_vregs(neg);
        if (!has_pos_v) return emit(VNEG, {neg_v});
        return emit(VSUB, {pos_v, neg_v});
    }

    // division ------------------------------------------------------------------------------------------------------
    VReg lower_div(int n) {
        const int num = g.nodes[n].args[0], den = g.nodes[n].args[1];
        double cn = 0, cd = 0;
        const bool hn = cv(num, &cn), hd = cv(den, &cd);
        if (hn && hd && cd != 0.0 && std::isfinite(cn / cd)) return add_const(cn / cd);
        const Node &dn = g.nodes[den];
        if (hn && std::fabs(cn - 1.0) < EPSILON && dn.kind == SUM && dn.args.size() == 2) { // 1/(exp(u)-1) (div.rs:19-27)
            const int pairs[2][2] = {{dn.args[0], dn.args[1]}, {dn.args[1], dn.args[0]}};
            for (auto &p : pairs) {
                const Node &a = g.nodes[p[0]], &b = g.nodes[p[1]];
                if (a.kind == CALL && a.name == "exp" && a.args.size() == 1 && b.kind == NUM && b.value == -1.0)
                    return emit(VRECIPEXPM1, {m[a.args[0]]});
            }
        }
        if (hd && cd != 0.0 && std::isfinite(cd)) return mul_vregs({m[num], add_const(1.0 / cd)}); // u/c -> u*(1/c)
        const Node &nn = g.nodes[num];
        if (nn.kind == CALL && nn.name == "sin" && nn.args.size() == 1 && nn.args[0] == den)
            return emit_x(VBUILTIN, {m[den]}, fnop_of("sinc"));
        if (hn && std::fabs(cn - 1.0) < EPSILON) return emit(VRECIP, {m[den]});
        return emit(VDIV, {m[num], m[den]});
    }

    // powers --------------------------------------------------------------------------------------------------------
    VReg lower_pow(int n) {
        const int base = g.nodes[n].args[0], ex = g.nodes[n].args[1];
        double cb = 0, ce = 0;
        const bool hb = cv(base, &cb), he = cv(ex, &ce);
        if (hb && he) {
            const double v = std::pow(cb, ce);
            if (std::isfinite(v)) return add_const(v);
        }
        const VReg b = m[base];
        if (hb && std::fabs(cb - 2.718281828459045) < EPSILON) return lower_exp(ex); // e^u is exp(u)
        if (he && std::isfinite(ce)) {
            if (ce == std::trunc(ce) && std::fabs(ce) < 2147483648.0) {
                const long long k = (long long)ce;
                if (k == 0) return add_const(1.0);
                if (k == 1) return b;
                switch (k) {
                case 2: return emit(VSQUARE, {b});
                case 3: return emit(VCUBE, {b});
                case 4: return emit(VPOW4, {b});
                case -1: return emit(VRECIP, {b});
                case -2: return emit(VINVSQUARE, {b});
                case -3: return emit(VINVCUBE, {b});
                default: return emit_x(VPOWI, {b}, k);
                }
            }
            if (ce == 0.5) return emit(VSQRT, {b});
            if (ce == -0.5) return emit(VINVSQRT, {b});
            if (ce == 1.5) return emit(VPOW3_2, {b});
            if (ce == -1.5) return emit(VINVPOW3_2, {b});
        }
        return emit(VPOW, {b, m[ex]});
    }

    // functions -----------------------------------------------------------------------------------------------------
    VReg lower_exp(int arg) {
        const Node &a = g.nodes[arg];
        auto is_two = [&](int e) { return g.nodes[e].kind == NUM && g.nodes[e].value == 2.0; };
        if (a.kind == POW && is_two(a.args[1])) return emit(VEXPSQR, {m[a.args[0]]}); // vir/matcher.rs:143-169
        if (a.kind == PROD && a.args.size() == 2) {
            const Node &a0 = g.nodes[a.args[0]], &a1 = g.nodes[a.args[1]];
            if (a0.kind == NUM && a0.value == -1.0 && a1.kind == POW && is_two(a1.args[1])) return emit(VEXPSQRNEG, {m[a1.args[0]]});
        }
        bool has_pos = false;
        VReg posv;
        if (a.kind == PROD) {
            has_pos = neg_inner(arg, &posv);
        } else if (a.kind == DIV) {
            const int num = a.args[0], den = a.args[1];
            double cd, cn;
            if (!cv(den, &cd)) {
                if (g.nodes[num].kind == PROD) {
                    VReg pn;
                    if (neg_inner(num, &pn)) {
                        posv = emit(VDIV, {pn, m[den]});
                        has_pos = true;
                    }
                } else if (cv(num, &cn) && cn < 0 && std::isfinite(cn)) {
                    posv = emit(VDIV, {add_const(-cn), m[den]});
                    has_pos = true;
                }
            }
        }
        if (has_pos) return emit_x(VBUILTIN, {posv}, fnop_of("exp_neg"));
        return emit(VEXP, {m[arg]});
    }
    bool lower_call(int n, VReg *out) {
        const Node &nd = g.nodes[n];
        std::string name = fn_alias(nd.name);
        const std::vector<int> &args = nd.args;
        if (args.size() == 1) {
            double c, v;
            if (cv(args[0], &c) && const_fold(name, c, &v)) {
                *out = add_const(v);
                return true;
            }
        }
        if (name == "exp" && args.size() == 1) {
            *out = lower_exp(args[0]);
            return true;
        }
        if (name == "log" && args.size() == 1) name = "ln";
        if ((name == "log2" || name == "log10") && args.size() == 1) {
            const VReg base = add_const(name == "log2" ? 2.0 : 10.0);
            *out = emit_x(VBUILTIN, {base, m[args[0]]}, fnop_of("log"));
            return true;
        }
        const int fn = fnop_of(name);
        if (fn < 0 || FNOPS[fn].arity != (int)args.size()) {
            err = "UnsupportedFunction: " + nd.name + "/" + std::to_string(args.size());
            return false;
        }
        if (name == "sin") { *out = emit(VSIN, {m[args[0]]}); return true; } // reg_alloc.rs:324-350
        if (name == "cos") { *out = emit(VCOS, {m[args[0]]}); return true; }
        if (name == "exp") { *out = emit(VEXP, {m[args[0]]}); return true; }
        if (name == "ln") { *out = emit(VLN, {m[args[0]]}); return true; }
        if (name == "sqrt") { *out = emit(VSQRT, {m[args[0]]}); return true; }
        std::vector<VReg> s;
        for (int a : args) s.push_back(m[a]);
        *out = emit_x(VBUILTIN, s, fn);
        return true;
    }

    bool lower_node(int n, VReg *out) {
        const Node &nd = g.nodes[n];
        switch (nd.kind) {
        case NUM: *out = add_const(nd.value); return true;
        case SYM: {
            auto it = param_idx.find(nd.name);
            if (it != param_idx.end()) { // a parameter named `pi` wins over the constant (misc.rs:14-26)
                *out = VReg{'p', it->second};
                return true;
            }
            double v;
            if (known_constant(nd.name, &v)) {
                *out = add_const(v);
                return true;
            }
            err = "UnboundVariable: " + nd.name;
            return false;
        }
        case SUM: *out = lower_sum(n); return true;
        case PROD: *out = lower_product(n); return true;
        case DIV: *out = lower_div(n); return true;
        case POW: *out = lower_pow(n); return true;
        case CALL: return lower_call(n, out);
        }
        err = "UnsupportedExpression";
        return false;
    }
    bool lower(int root, VReg *out) {
        for (int n : g.postorder(root)) {
            if (have[n]) continue;
            VReg r;
            if (!lower_node(n, &r)) return false;
            m[n] = r;
            have[n] = 1;
        }
        *out = m[root];
        return true;
    }
};

// Sin(a) ... Cos(a) -> SinCos (fusion.rs:105-187, non-adjacent)
std::vector<VInstr> fuse_sincos(const std::vector<VInstr> &instrs) {
    std::map<std::pair<int, VReg>, size_t> first; // (op, source) -> index
    std::vector<VInstr> out = instrs;
    std::set<size_t> dead;
    for (size_t idx = 0; idx < out.size(); ++idx) {
        const VOp op = out[idx].op;
        if (op != VSIN && op != VCOS) continue;
        const VReg src = out[idx].srcs[0];
        const auto other = std::make_pair((int)(op == VSIN ? VCOS : VSIN), src);
        auto it = first.find(other);
        if (it != first.end() && !dead.count(it->second) && (out[it->second].op == VSIN || out[it->second].op == VCOS)) {
            const size_t j = it->second;
            const VReg sin_d = out[j].op == VSIN ? out[j].dests[0] : out[idx].dests[0];
            const VReg cos_d = out[j].op == VSIN ? out[idx].dests[0] : out[j].dests[0];
            VInstr sc;
            sc.op = VSINCOS;
            sc.dests = {sin_d, cos_d};
            sc.srcs = {src};
            out[j] = sc;
            dead.insert(idx);
            continue;
        }
        first.emplace(std::make_pair((int)op, src), idx);
    }
    std::vector<VInstr> kept;
    for (size_t k = 0; k < out.size(); ++k)
        if (!dead.count(k)) kept.push_back(out[k]);
    return kept;
}

std::vector<VInstr> dce(const std::vector<VInstr> &instrs, const std::vector<VReg> &roots) {
    std::set<VReg> live;
    for (const VReg &r : roots)
        if (r.kind == 't') live.insert(r);
    std::vector<VInstr> keep;
    for (size_t i = instrs.size(); i-- > 0;) {
        bool any = false;
        for (const VReg &d : instrs[i].dests) any |= live.count(d) != 0;
        if (!any) continue;
        keep.push_back(instrs[i]);
        for (const VReg &s : instrs[i].srcs)
            if (s.kind == 't') live.insert(s);
    }
    return std::vector<VInstr>(keep.rbegin(), keep.rend());
}

uint32_t ref_opcode(VOp op) {
    switch (op) {
    case VNEG: return R_NEG;
    case VNEGMUL: return R_NEGMUL;
    case VMULADD: return R_MULADD;
    case VMULSUB: return R_MULSUB;
    case VNEGMULADD: return R_NEGMULADD;
    case VSUB: return R_SUB;
    case VDIV: return R_DIV;
    case VRECIP: return R_RECIP;
    case VSQUARE: return R_SQUARE;
    case VCUBE: return R_CUBE;
    case VPOW4: return R_POW4;
    case VINVSQUARE: return R_INVSQUARE;
    case VINVCUBE: return R_INVCUBE;
    case VPOWI: return R_POWI;
    case VSQRT: return R_SQRT;
    case VINVSQRT: return R_INVSQRT;
    case VPOW3_2: return R_POW3_2;
    case VINVPOW3_2: return R_INVPOW3_2;
    case VPOW: return R_POW;
    case VEXP: return R_EXP;
    case VEXPSQR: return R_EXPSQR;
    case VEXPSQRNEG: return R_EXPSQRNEG;
    case VSIN: return R_SIN;
    case VCOS: return R_COS;
    case VLN: return R_LN;
    case VRECIPEXPM1: return R_RECIPEXPM1;
    case VSINCOS: return R_SINCOS;
    default: return R_END;
    }
}

} // namespace

// Compiles `exprs` over `params` into ONE multi-output program in the reference's wire format.
int compile(const std::vector<std::string> &exprs, const std::vector<std::string> &params, RefProgram &out, std::string &err) {
    Dag dag;
    std::vector<int> roots_e;
    for (const std::string &src : exprs) {
        Parser p(dag);
        const int e = p.parse(src);
        if (e < 0) {
            err = "parse error: " + p.err;
            return SAF_ERR_INVALID_ARGUMENT;
        }
        roots_e.push_back(e);
    }
    Gen gen(dag, params);
    std::vector<VReg> roots;
    for (int e : roots_e) {
        VReg r;
        if (!gen.lower(e, &r)) {
            err = gen.err;
            return SAF_ERR_INVALID_PROGRAM;
        }
        roots.push_back(r);
    }
    const std::vector<VInstr> instrs = dce(fuse_sincos(gen.instrs), roots);

    // constant compaction: only referenced constants survive, in order of first reference (compact.rs:10-62)
    std::map<uint32_t, uint32_t> used_c;
    for (const VInstr &in : instrs)
        for (const VReg &s : in.srcs)
            if (s.kind == 'c' && !used_c.count(s.idx)) {
                const uint32_t k = (uint32_t)used_c.size();
                used_c[s.idx] = k;
            }
    for (const VReg &r : roots)
        if (r.kind == 'c' && !used_c.count(r.idx)) {
            const uint32_t k = (uint32_t)used_c.size();
            used_c[r.idx] = k;
        }
    out = RefProgram();
    out.consts.assign(used_c.size(), 0.0);
    for (auto &kv : used_c) out.consts[kv.second] = gen.consts[kv.first];
    const uint32_t n_params = (uint32_t)params.size();
    const uint32_t temp_base = n_params + (uint32_t)out.consts.size();

    // linear scan over the temporaries, LIFO free list (reg_alloc.rs:106-138,441-449)
    std::map<VReg, size_t> last_use;
    for (size_t idx = 0; idx < instrs.size(); ++idx)
        for (const VReg &s : instrs[idx].srcs)
            if (s.kind == 't') last_use[s] = idx;
    for (const VReg &r : roots)
        if (r.kind == 't') last_use[r] = instrs.size();
    std::map<VReg, uint32_t> phys;
    std::vector<uint32_t> free_list;
    uint32_t n_phys = 0;
    auto reg = [&](const VReg &v) -> uint32_t {
        if (v.kind == 'p') return v.idx;
        if (v.kind == 'c') return n_params + used_c[v.idx];
        return temp_base + phys[v];
    };
    std::vector<uint32_t> &flat = out.flat;
    for (size_t idx = 0; idx < instrs.size(); ++idx) {
        const VInstr &in = instrs[idx];
        for (const VReg &d : in.dests) {
            if (!free_list.empty()) {
                phys[d] = free_list.back();
                free_list.pop_back();
            } else
                phys[d] = n_phys++;
        }
        std::vector<uint32_t> s, d;
        for (const VReg &x : in.srcs) s.push_back(reg(x));
        for (const VReg &x : in.dests) d.push_back(reg(x));
        if (in.op == VADD || in.op == VMUL) {
            const bool add = in.op == VADD;
            if (s.size() <= 4) {
                flat.push_back(s.size() == 2 ? (add ? R_ADD : R_MUL) : s.size() == 3 ? (add ? R_ADD3 : R_MUL3) : (add ? R_ADD4 : R_MUL4));
                flat.push_back(d[0]);
                for (uint32_t x : s) flat.push_back(x);
            } else {
                flat.push_back(add ? R_ADDN : R_MULN);
                flat.push_back(d[0]);
                flat.push_back((uint32_t)out.pool.size());
                flat.push_back((uint32_t)s.size());
                out.pool.insert(out.pool.end(), s.begin(), s.end());
            }
        } else if (in.op == VSINCOS) {
            flat.insert(flat.end(), {(uint32_t)R_SINCOS, d[0], d[1], s[0]});
        } else if (in.op == VPOWI) {
            flat.insert(flat.end(), {(uint32_t)R_POWI, d[0], s[0], (uint32_t)(int32_t)in.extra});
        } else if (in.op == VBUILTIN) {
            flat.push_back((uint32_t)(R_BUILTIN1 + (s.size() - 1)));
            flat.push_back(d[0]);
            flat.push_back((uint32_t)in.extra);
            for (uint32_t x : s) flat.push_back(x);
        } else {
            flat.push_back(ref_opcode(in.op));
            flat.push_back(d[0]);
            for (uint32_t x : s) flat.push_back(x);
        }
        // sources die after the instruction is emitted, in order of first occurrence
        std::vector<VReg> seen;
        for (const VReg &x : in.srcs) {
            bool dup = false;
            for (const VReg &y : seen) dup |= y == x;
            if (dup) continue;
            seen.push_back(x);
            auto it = last_use.find(x);
            if (x.kind == 't' && it != last_use.end() && it->second == idx) free_list.push_back(phys[x]);
        }
        for (const VReg &x : in.dests) // a destination nobody reads (one half of a SinCos)
            if (!last_use.count(x)) free_list.push_back(phys[x]);
    }
    flat.push_back(R_END);
    out.param_count = n_params;
    out.workspace_size = temp_base + n_phys;
    for (const VReg &r : roots) out.result_regs.push_back(reg(r));
    return SAF_OK;
}

} // namespace cc
} // namespace saf
