"""Expression -> SymbAnaFis bytecode compiler (host side, not on the hot path).

The Rust reference compiles with a 7 kLoC pipeline (``src/evaluator/logic/bytecode/compile/**``,
SURVEY.md section 2 row 6).  No Rust toolchain exists in this image, so programs for the evaluator are
produced here.  This is a *re-implementation of the documented lowering rules*, not a byte-for-byte
clone: it emits valid programs in the reference ISA using the same instruction-selection ideas

* sums      -> MulAdd / MulSub / NegMulAdd / Sub / Add / Add3 / Add4 / AddN   (``codegen/lower/sum.rs:104-344``)
* products  -> constants folded into one trailing factor, Mul..MulN          (``codegen/lower/product.rs:10-91``)
* division  -> RecipExpm1, ``u/c -> u*(1/c)``, Sinc, Div                       (``codegen/lower/div.rs:13-68``)
* powers    -> Square/Cube/Pow4/Recip/InvSquare/InvCube/Powi/Sqrt/InvSqrt/Pow3_2/InvPow3_2/Pow
                                                                              (``codegen/lower/pow.rs:186-334``)
* functions -> ExpSqr/ExpSqrNeg, Builtin1{ExpNeg}, FnOp table, log2/log10     (``codegen/lower/func.rs:14-110``)
* Sin+Cos of one argument -> SinCos                                           (``optimize/fusion.rs:105-187``)
* linear-scan allocation with a LIFO free list, sources released after the instruction so a fresh
  destination never aliases a dying source                                   (``emit/reg_alloc.rs:106-138,441-449``)
* unified register layout params | consts | temps                             (``instruction.rs:18-21``)

Several expressions over the same parameters compile into ONE program with several result
registers and shared subexpressions (the fused multi-output programs of the north star).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .expr import KNOWN_CONSTANTS, Expr, postorder
from .isa import FNOP, FNOP_ARITY, Program, assemble

EPSILON = 1e-14  # crate::EPSILON, src/lib.rs:321

VReg = Tuple[str, int]  # ("p", i) | ("c", i) | ("t", i)


class CompileError(RuntimeError):
    """UnboundVariable / UnsupportedFunction / UnsupportedExpression (``api.rs:402-408``)."""


# name -> FnOp name (``vir/registry.rs:14-87``)
_FN_ALIASES = {
    "sign": "signum", "sgn": "signum", "ynm": "spherical_harmonic", "lambertw": "lambert_w",
    "besselj": "bessel_j", "bessely": "bessel_y", "besseli": "bessel_i", "besselk": "bessel_k",
    "ellipk": "elliptic_k", "ellipe": "elliptic_e", "loggamma": "lgamma", "exp_neg": "exp_neg",
    "assoc_legendre": "assoc_legendre", "spherical_harmonic": "spherical_harmonic",
    "zeta_deriv": "zeta_deriv",
}

# constant folding of unary functions (``vir/registry.rs:89-160``)
_CONST_FOLD = {
    "sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp, "sqrt": math.sqrt,
    "ln": math.log, "abs": abs, "sinh": math.sinh, "cosh": math.cosh, "tanh": math.tanh,
    "asin": math.asin, "acos": math.acos, "atan": math.atan,
}


class _Gen:
    def __init__(self, params: Sequence[str]):
        self.params = list(params)
        self.param_idx = {n: i for i, n in reversed(list(enumerate(self.params)))}
        self.consts: List[float] = []
        self.const_idx: Dict[bytes, int] = {}
        self.instrs: List[list] = []  # [op, dests(list), srcs(list), extra]
        self.n_temps = 0
        self.gvn: Dict[tuple, VReg] = {}

    # -- plumbing ----------------------------------------------------------------------------
    def add_const(self, v: float) -> VReg:
        key = np.float64(v).tobytes()  # interned by bit pattern (compiler.rs:61-73)
        i = self.const_idx.get(key)
        if i is None:
            i = len(self.consts)
            self.consts.append(float(v))
            self.const_idx[key] = i
        return ("c", i)

    def const_val(self, r: VReg) -> Optional[float]:
        return self.consts[r[1]] if r[0] == "c" else None

    def emit(self, op: str, srcs: Sequence[VReg], extra=None, n_dest: int = 1):
        srcs = list(srcs)
        if op in ("Add", "Mul") and len(srcs) == 2:
            key_srcs = tuple(sorted(srcs))  # commutative
        else:
            key_srcs = tuple(srcs)
        key = (op, key_srcs, extra)
        hit = self.gvn.get(key)
        if hit is not None:
            return hit
        dests = [("t", self.n_temps + k) for k in range(n_dest)]
        self.n_temps += n_dest
        self.instrs.append([op, dests, srcs, extra])
        res = dests[0] if n_dest == 1 else tuple(dests)
        self.gvn[key] = res
        return res

    def add_vregs(self, v: List[VReg]) -> VReg:
        if not v:
            return self.add_const(0.0)
        if len(v) == 1:
            return v[0]
        return self.emit("Add", v)

    def mul_vregs(self, v: List[VReg]) -> VReg:
        if not v:
            return self.add_const(1.0)
        if len(v) == 1:
            return v[0]
        if len(v) == 2 and v[0] == v[1]:
            return self.emit("Square", [v[0]])
        return self.emit("Mul", v)

    # -- node lowering -----------------------------------------------------------------------
    def lower(self, root: Expr, memo: Dict[int, VReg]) -> VReg:
        for n in postorder(root):
            if id(n) in memo:
                continue
            memo[id(n)] = self.lower_node(n, memo)
        return memo[id(root)]

    def lower_node(self, n: Expr, m: Dict[int, VReg]) -> VReg:
        k = n.kind
        if k == "num":
            return self.add_const(n.value)
        if k == "sym":
            if n.name in self.param_idx:  # a parameter named `pi` wins over the constant (misc.rs:14-26)
                return ("p", self.param_idx[n.name])
            if n.name in KNOWN_CONSTANTS:
                return self.add_const(KNOWN_CONSTANTS[n.name])
            raise CompileError(f"UnboundVariable: {n.name}")
        if k == "sum":
            return self.lower_sum(n, m)
        if k == "prod":
            return self.lower_product(n, m)
        if k == "div":
            return self.lower_div(n, m)
        if k == "pow":
            return self.lower_pow(n, m)
        if k == "call":
            return self.lower_call(n, m)
        raise CompileError(f"UnsupportedExpression: {k}")

    def cv(self, e: Expr, m) -> Optional[float]:
        return self.const_val(m[id(e)])

    # products -------------------------------------------------------------------------------
    def _split_product(self, e: Expr, m) -> Tuple[float, bool, List[VReg]]:
        """(constant part, has constant, variable vregs) of a product node."""
        c, has, v = 1.0, False, []
        for f in e.args:
            fc = self.cv(f, m)
            if fc is not None:
                c *= fc
                has = True
            else:
                v.append(m[id(f)])
        return c, has, v

    def lower_product(self, n: Expr, m) -> VReg:
        c, _has, v = self._split_product(n, m)
        if math.isfinite(c):
            if c == 0.0:
                c = 0.0  # -0.0 -> +0.0 (product.rs:67-71)
            if abs(c - 1.0) > EPSILON:
                if not v:
                    return self.add_const(c)
                if abs(c + 1.0) < EPSILON and len(v) == 1:
                    return self.emit("Neg", v)
                if abs(c + 1.0) < EPSILON and len(v) == 2:
                    return self.emit("NegMul", v)
                v = v + [self.add_const(c)]  # constant goes last (product.rs:62-73)
        else:
            v = v + [self.add_const(c)]
        return self.mul_vregs(v)

    def _neg_inner(self, e: Expr, m) -> Optional[VReg]:
        """``try_extract_negated_product`` (sum.rs:36-68): vreg of e without its -1 factor."""
        if e.kind != "prod":
            return None
        c, has, v = self._split_product(e, m)
        if not has or not math.isfinite(c) or c >= 0.0:
            return None
        pos = -c
        if abs(pos - 1.0) > EPSILON:
            v = v + [self.add_const(pos)]
        return self.mul_vregs(v)

    def _prod2(self, e: Expr, m) -> Optional[Tuple[VReg, VReg]]:
        """A product that is exactly two register factors (after constant folding)."""
        if e.kind != "prod":
            return None
        c, has, v = self._split_product(e, m)
        if not math.isfinite(c):
            return None
        if not has or abs(c - 1.0) <= EPSILON:
            return (v[0], v[1]) if len(v) == 2 else None
        if c > 0 and len(v) == 1:
            return (v[0], self.add_const(c))
        return None

    def _negprod2(self, e: Expr, m) -> Optional[Tuple[VReg, VReg]]:
        """``-(a*b)`` as two register factors."""
        if e.kind != "prod":
            return None
        c, has, v = self._split_product(e, m)
        if not has or not math.isfinite(c) or c >= 0.0:
            return None
        if abs(c + 1.0) <= EPSILON:
            return (v[0], v[1]) if len(v) == 2 else None
        if len(v) == 1:
            return (v[0], self.add_const(-c))
        return None

    # sums -----------------------------------------------------------------------------------
    def lower_sum(self, n: Expr, m) -> VReg:
        terms = n.args
        if len(terms) == 2:
            t0, t1 = terms
            c0, c1 = self.cv(t0, m), self.cv(t1, m)
            if c0 is not None and c1 is not None and math.isfinite(c0 + c1):
                return self.add_const(c0 + c1)
            if c0 is not None and math.isfinite(c0) and abs(c0) < EPSILON:
                return m[id(t1)]
            if c1 is not None and math.isfinite(c1) and abs(c1) < EPSILON:
                return m[id(t0)]
            for a, b in ((t0, t1), (t1, t0)):  # a*b - c  -> MulSub
                p = self._prod2(a, m)
                if p is not None and c0 is None and c1 is None:
                    ng = self._neg_inner(b, m)
                    if ng is not None:
                        return self.emit("MulSub", [p[0], p[1], ng])
            for a, b in ((t0, t1), (t1, t0)):  # c - a*b  -> NegMulAdd
                p = self._negprod2(a, m)
                if p is not None:
                    return self.emit("NegMulAdd", [p[0], p[1], m[id(b)]])
            for a, b in ((t0, t1), (t1, t0)):  # a*b + c  -> MulAdd
                p = self._prod2(a, m)
                if p is not None:
                    return self.emit("MulAdd", [p[0], p[1], m[id(b)]])
            for a, b in ((t0, t1), (t1, t0)):  # x - y    -> Sub
                ng = self._neg_inner(a, m)
                if ng is not None:
                    return self.emit("Sub", [m[id(b)], ng])
            return self.emit("Add", [m[id(t0)], m[id(t1)]])

        pos: List[VReg] = []
        neg: List[VReg] = []
        cacc, has_c = 0.0, False
        pp = np_ = None
        for t in terms:
            c = self.cv(t, m)
            if c is not None:
                cacc += c
                has_c = True
                continue
            q = self._negprod2(t, m)
            if q is not None:
                if np_ is None:
                    np_ = q
                else:
                    neg.append(self.mul_vregs(list(q)))
                continue
            q = self._prod2(t, m)
            if q is not None:
                if pp is None:
                    pp = q
                else:
                    pos.append(self.mul_vregs(list(q)))
                continue
            ng = self._neg_inner(t, m)
            if ng is not None:
                neg.append(ng)
            else:
                pos.append(m[id(t)])
        if has_c:
            if math.isfinite(cacc):
                if abs(cacc) > EPSILON:
                    (pos if cacc > 0 else neg).append(self.add_const(abs(cacc)))
            else:
                pos.append(self.add_const(cacc))
        return self._assemble_sum(pp, np_, pos, neg)

    def _assemble_sum(self, pp, np_, pos: List[VReg], neg: List[VReg]) -> VReg:
        """``assemble_nary_sum`` (sum.rs:261-344)."""
        if pp is not None and np_ is None and not neg:
            if not pos:
                return self.mul_vregs(list(pp))
            return self.emit("MulAdd", [pp[0], pp[1], self.add_vregs(pos)])
        if pp is None and np_ is None and not neg:
            return self.add_vregs(pos)
        if pp is None and np_ is not None and not pos:
            inner = self.mul_vregs(list(np_)) if not neg else self.add_vregs(neg + [self.mul_vregs(list(np_))])
            return self.emit("Neg", [inner])
        if pp is None and np_ is None and not pos:
            return self.emit("Neg", [self.add_vregs(neg)])
        if pp is not None and np_ is None and not pos:
            return self.emit("MulSub", [pp[0], pp[1], self.add_vregs(neg)])
        if pp is None and np_ is not None and not neg:
            return self.emit("NegMulAdd", [np_[0], np_[1], self.add_vregs(pos)])
        # both sides present
        if pp is not None:
            pos_v = self.mul_vregs(list(pp)) if not pos else self.emit("MulAdd", [pp[0], pp[1], self.add_vregs(pos)])
        else:
            pos_v = self.add_vregs(pos) if pos else None
        if np_ is not None:
            neg = neg + [self.mul_vregs(list(np_))]
        neg_v = self.add_vregs(neg)
        if pos_v is None:
            return self.emit("Neg", [neg_v])
        return self.emit("Sub", [pos_v, neg_v])

    # division -------------------------------------------------------------------------------
    def lower_div(self, n: Expr, m) -> VReg:
        num, den = n.args
        cn, cd = self.cv(num, m), self.cv(den, m)
        if cn is not None and cd is not None and cd != 0.0 and math.isfinite(cn / cd):
            return self.add_const(cn / cd)
        # 1/(exp(u)-1) -> RecipExpm1 (div.rs:19-27)
        if cn is not None and abs(cn - 1.0) < EPSILON and den.kind == "sum" and len(den.args) == 2:
            for a, b in (den.args, den.args[::-1]):
                if a.kind == "call" and a.name == "exp" and len(a.args) == 1 and b.kind == "num" and b.value == -1.0:
                    return self.emit("RecipExpm1", [m[id(a.args[0])]])
        if cd is not None and cd != 0.0 and math.isfinite(cd):  # u/c -> u*(1/c) (div.rs:29-44)
            return self.mul_vregs([m[id(num)], self.add_const(1.0 / cd)])
        if num.kind == "call" and num.name == "sin" and len(num.args) == 1 and num.args[0] is den:
            return self.emit("Builtin1", [m[id(den)]], extra=FNOP["sinc"])
        if cn is not None and abs(cn - 1.0) < EPSILON:
            return self.emit("Recip", [m[id(den)]])
        return self.emit("Div", [m[id(num)], m[id(den)]])

    # powers ---------------------------------------------------------------------------------
    def lower_pow(self, n: Expr, m) -> VReg:
        base, ex = n.args
        cb, ce = self.cv(base, m), self.cv(ex, m)
        if cb is not None and ce is not None:
            try:
                v = math.pow(cb, ce)
                if math.isfinite(v):
                    return self.add_const(v)
            except (OverflowError, ValueError):
                pass
        b = m[id(base)]
        if cb is not None and abs(cb - math.e) < EPSILON:  # e^u is exp(u) (pow.rs)
            return self._lower_exp(ex, m)
        if ce is not None and math.isfinite(ce):
            if ce == int(ce) and abs(ce) < 2 ** 31:
                k = int(ce)
                simple = {0: None, 1: None, 2: "Square", 3: "Cube", 4: "Pow4", -1: "Recip",
                          -2: "InvSquare", -3: "InvCube"}
                if k == 0:
                    return self.add_const(1.0)
                if k == 1:
                    return b
                if k in simple:
                    return self.emit(simple[k], [b])
                return self.emit("Powi", [b], extra=k)
            frac = {0.5: "Sqrt", -0.5: "InvSqrt", 1.5: "Pow3_2", -1.5: "InvPow3_2"}
            if ce in frac:
                return self.emit(frac[ce], [b])
        return self.emit("Pow", [b, m[id(ex)]])

    # functions ------------------------------------------------------------------------------
    def _lower_exp(self, arg: Expr, m) -> VReg:
        # exp(+-u^2) -> ExpSqr / ExpSqrNeg (vir/matcher.rs:143-169)
        if arg.kind == "pow" and arg.args[1].kind == "num" and arg.args[1].value == 2.0:
            return self.emit("ExpSqr", [m[id(arg.args[0])]])
        if arg.kind == "prod" and len(arg.args) == 2:
            a0, a1 = arg.args
            if a0.kind == "num" and a0.value == -1.0 and a1.kind == "pow" and a1.args[1].kind == "num" \
                    and a1.args[1].value == 2.0:
                return self.emit("ExpSqrNeg", [m[id(a1.args[0])]])
        # exp(-c * ...) and exp(-.../d) -> Builtin1{ExpNeg} on the positive part (exp_helpers.rs:30-108)
        posv = None
        if arg.kind == "prod":
            posv = self._neg_inner(arg, m)
        elif arg.kind == "div":
            num, den = arg.args
            if self.cv(den, m) is None:
                if num.kind == "prod":
                    pn = self._neg_inner(num, m)
                    if pn is not None:
                        posv = self.emit("Div", [pn, m[id(den)]])
                else:
                    cn = self.cv(num, m)
                    if cn is not None and cn < 0 and math.isfinite(cn):
                        posv = self.emit("Div", [self.add_const(-cn), m[id(den)]])
        if posv is not None:
            return self.emit("Builtin1", [posv], extra=FNOP["exp_neg"])
        return self.emit("Exp", [m[id(arg)]])

    def lower_call(self, n: Expr, m) -> VReg:
        name = _FN_ALIASES.get(n.name, n.name)
        args = n.args
        if len(args) == 1:
            c = self.cv(args[0], m)
            if c is not None and name in _CONST_FOLD:
                try:
                    v = _CONST_FOLD[name](c)
                    if math.isfinite(v):
                        return self.add_const(v)
                except (ValueError, OverflowError):
                    pass
        if name == "exp" and len(args) == 1:
            return self._lower_exp(args[0], m)
        if name == "log" and len(args) == 1:
            name = "ln"
        if name in ("log2", "log10") and len(args) == 1:
            base = self.add_const(2.0 if name == "log2" else 10.0)
            return self.emit("Builtin2", [base, m[id(args[0])]], extra=FNOP["log"])
        if name not in FNOP or FNOP_ARITY[FNOP[name]] != len(args):
            raise CompileError(f"UnsupportedFunction: {n.name}/{len(args)}")
        direct = {"sin": "Sin", "cos": "Cos", "exp": "Exp", "ln": "Ln", "sqrt": "Sqrt"}
        if name in direct:  # reg_alloc.rs:324-350
            return self.emit(direct[name], [m[id(args[0])]])
        return self.emit(f"Builtin{len(args)}", [m[id(a)] for a in args], extra=FNOP[name])


def _fuse_sincos(instrs: List[list]) -> List[list]:
    """Sin(a) ... Cos(a) -> SinCos (fusion.rs:105-187, non-adjacent)."""
    first: Dict[Tuple[str, VReg], int] = {}
    out = [list(i) for i in instrs]
    dead = set()
    for idx, (op, dests, srcs, _extra) in enumerate(out):
        if op not in ("Sin", "Cos"):
            continue
        other = ("Cos" if op == "Sin" else "Sin", srcs[0])
        j = first.get(other)
        if j is not None and j not in dead and out[j][0] in ("Sin", "Cos"):
            sin_d = out[j][1][0] if out[j][0] == "Sin" else dests[0]
            cos_d = dests[0] if out[j][0] == "Sin" else out[j][1][0]
            out[j] = ["SinCos", [sin_d, cos_d], [srcs[0]], None]
            dead.add(idx)
            continue
        first.setdefault((op, srcs[0]), idx)
    return [ins for k, ins in enumerate(out) if k not in dead]


def _dce(instrs: List[list], roots: Sequence[VReg]) -> List[list]:
    live = {r for r in roots if r[0] == "t"}
    keep = []
    for ins in reversed(instrs):
        if any(d in live for d in ins[1]):
            keep.append(ins)
            for s in ins[2]:
                if s[0] == "t":
                    live.add(s)
    keep.reverse()
    return keep


def compile_exprs(exprs: Sequence[Expr], params: Sequence[str]) -> Program:
    """Compile one or more expressions over ``params`` into a single (multi-output) program."""
    gen = _Gen(params)
    memo: Dict[int, VReg] = {}
    roots = [gen.lower(e, memo) for e in exprs]
    instrs = _dce(_fuse_sincos(gen.instrs), roots)

    # constant compaction: only referenced constants survive (compact.rs:10-62)
    used_c: Dict[int, int] = {}
    for ins in instrs:
        for s in ins[2]:
            if s[0] == "c" and s[1] not in used_c:
                used_c[s[1]] = len(used_c)
    for r in roots:
        if r[0] == "c" and r[1] not in used_c:
            used_c[r[1]] = len(used_c)
    consts = [0.0] * len(used_c)
    for old, new in used_c.items():
        consts[new] = gen.consts[old]
    n_params = len(params)
    temp_base = n_params + len(consts)

    # linear scan over temps
    last_use: Dict[VReg, int] = {}
    for idx, ins in enumerate(instrs):
        for s in ins[2]:
            if s[0] == "t":
                last_use[s] = idx
    for r in roots:
        if r[0] == "t":
            last_use[r] = len(instrs)
    phys: Dict[VReg, int] = {}
    free: List[int] = []
    n_phys = 0
    arg_pool: List[int] = []
    out: List[tuple] = []

    def reg(v: VReg) -> int:
        if v[0] == "p":
            return v[1]
        if v[0] == "c":
            return n_params + used_c[v[1]]
        return temp_base + phys[v]

    for idx, (op, dests, srcs, extra) in enumerate(instrs):
        for d in dests:
            if free:
                phys[d] = free.pop()
            else:
                phys[d] = n_phys
                n_phys += 1
        s = [reg(x) for x in srcs]
        d = [reg(x) for x in dests]
        if op in ("Add", "Mul"):
            if len(s) <= 4:
                name = op if len(s) == 2 else f"{op}{len(s)}"
                out.append((name, d[0], *s))
            else:
                start = len(arg_pool)
                arg_pool.extend(s)
                out.append((f"{op}N", d[0], start, len(s)))
        elif op == "SinCos":
            out.append(("SinCos", d[0], d[1], s[0]))
        elif op == "Powi":
            out.append(("Powi", d[0], s[0], extra))
        elif op.startswith("Builtin"):
            out.append((op, d[0], extra, *s))
        else:
            out.append((op, d[0], *s))
        for x in dict.fromkeys(srcs):  # sources die after the instruction is emitted (in order of first occurrence)
            if x[0] == "t" and last_use.get(x) == idx:
                free.append(phys[x])
        for x in dests:  # a destination nobody reads (one half of a SinCos)
            if x not in last_use:
                free.append(phys[x])

    result_regs = [reg(r) for r in roots]
    return Program(
        flat=assemble(out),
        consts=np.asarray(consts, dtype=np.float64),
        arg_pool=np.asarray(arg_pool, dtype=np.uint32),
        param_count=n_params,
        workspace_size=temp_base + n_phys,
        result_regs=result_regs,
        param_names=list(params),
    )


def compile_expr(expr: Union[Expr, str], params: Sequence[str]) -> Program:
    from .expr import parse
    if isinstance(expr, str):
        expr = parse(expr)
    return compile_exprs([expr], params)
