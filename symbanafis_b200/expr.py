"""Minimal symbolic front end: hash-consed expression DAG, Pratt parser, d/dx.

Host-side only (microsecond-scale work, never on the hot path).  The Rust reference keeps a
much larger symbolic layer (``src/core``, ``src/parser``, ``src/diff``, ``src/simplification``,
SURVEY.md section 2 rows 11-14, out of scope); this module is the small subset needed to *produce
programs* for the evaluator when no Rust toolchain is available: the node kinds of
``src/core/expr/api.rs:51-127`` (Number, Symbol, FunctionCall, Sum, Product, Div, Pow), the
parser's conventions (``a - b`` is ``Sum[a, Product[-1, b]]``, ``src/parser/logic/pratt.rs:108``)
and textbook differentiation rules.
"""
from __future__ import annotations

import math
import re
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

Number = Union[int, float]

KNOWN_CONSTANTS = {"pi": math.pi, "PI": math.pi, "Pi": math.pi, "e": math.e, "E": math.e}

_INTERN: Dict[tuple, "Expr"] = {}


class Expr:
    """Immutable, interned expression node.  ``kind`` in num/sym/call/sum/prod/div/pow."""

    __slots__ = ("kind", "value", "name", "args", "_key", "__weakref__")

    def __new__(cls, kind: str, value=None, name: Optional[str] = None, args: Tuple["Expr", ...] = ()):
        if kind == "num":
            v = float(value)
            key = ("num", v.hex() if v == v else "nan")
        else:
            key = (kind, name, tuple(id(a) for a in args))
        node = _INTERN.get(key)
        if node is not None:
            return node
        node = object.__new__(cls)
        node.kind = kind
        node.value = float(value) if kind == "num" else None
        node.name = name
        node.args = tuple(args)
        node._key = key
        _INTERN[key] = node
        return node

    # -- constructors -----------------------------------------------------------------------
    @staticmethod
    def number(v: Number) -> "Expr":
        return Expr("num", value=v)

    @staticmethod
    def symbol(name: str) -> "Expr":
        return Expr("sym", name=name)

    @staticmethod
    def call(name: str, *args: "ExprLike") -> "Expr":
        return Expr("call", name=name, args=tuple(as_expr(a) for a in args))

    @staticmethod
    def sum(terms: Iterable["ExprLike"]) -> "Expr":
        flat: List[Expr] = []
        const = 0.0
        has_const = False
        for t in terms:
            t = as_expr(t)
            if t.kind == "sum":
                inner = t.args
            else:
                inner = (t,)
            for u in inner:
                if u.kind == "num":
                    const += u.value
                    has_const = True
                else:
                    flat.append(u)
        if has_const and (const != 0.0 or not flat):
            flat.append(Expr.number(const))
        if not flat:
            return Expr.number(0.0)
        if len(flat) == 1:
            return flat[0]
        return Expr("sum", args=tuple(flat))

    @staticmethod
    def product(factors: Iterable["ExprLike"]) -> "Expr":
        flat: List[Expr] = []
        coeff = 1.0
        for f in factors:
            f = as_expr(f)
            inner = f.args if f.kind == "prod" else (f,)
            for u in inner:
                if u.kind == "num":
                    coeff *= u.value
                else:
                    flat.append(u)
        if coeff == 0.0:
            return Expr.number(0.0)
        if not flat:
            return Expr.number(coeff)
        if coeff != 1.0:
            flat.insert(0, Expr.number(coeff))
        if len(flat) == 1:
            return flat[0]
        return Expr("prod", args=tuple(flat))

    @staticmethod
    def div(a: "ExprLike", b: "ExprLike") -> "Expr":
        a, b = as_expr(a), as_expr(b)
        if b.kind == "num" and b.value == 1.0:
            return a
        if a.kind == "num" and b.kind == "num" and b.value != 0.0:
            return Expr.number(a.value / b.value)
        if a.kind == "num" and a.value == 0.0:
            return a
        return Expr("div", args=(a, b))

    @staticmethod
    def pow(a: "ExprLike", b: "ExprLike") -> "Expr":
        a, b = as_expr(a), as_expr(b)
        if b.kind == "num":
            if b.value == 1.0:
                return a
            if b.value == 0.0:
                return Expr.number(1.0)
            if a.kind == "num":
                try:
                    v = math.pow(a.value, b.value)
                    if math.isfinite(v):
                        return Expr.number(v)
                except (OverflowError, ValueError):
                    pass
        return Expr("pow", args=(a, b))

    # -- operators --------------------------------------------------------------------------
    def __add__(self, o): return Expr.sum([self, o])
    def __radd__(self, o): return Expr.sum([o, self])
    def __sub__(self, o): return Expr.sum([self, Expr.product([-1.0, o])])
    def __rsub__(self, o): return Expr.sum([o, Expr.product([-1.0, self])])
    def __mul__(self, o): return Expr.product([self, o])
    def __rmul__(self, o): return Expr.product([o, self])
    def __truediv__(self, o): return Expr.div(self, o)
    def __rtruediv__(self, o): return Expr.div(o, self)
    def __pow__(self, o): return Expr.pow(self, o)
    def __rpow__(self, o): return Expr.pow(o, self)
    def __neg__(self): return Expr.product([-1.0, self])
    def __pos__(self): return self

    def __hash__(self): return id(self)
    def __eq__(self, o): return self is o

    # -- convenience ------------------------------------------------------------------------
    def sin(self): return Expr.call("sin", self)
    def cos(self): return Expr.call("cos", self)
    def exp(self): return Expr.call("exp", self)
    def ln(self): return Expr.call("ln", self)
    def sqrt(self): return Expr.call("sqrt", self)

    def diff(self, var: str) -> "Expr":
        return diff(self, var)

    def variables(self) -> List[str]:
        seen, out, stack = set(), [], [self]
        while stack:
            n = stack.pop()
            if id(n) in seen:
                continue
            seen.add(id(n))
            if n.kind == "sym":
                out.append(n.name)
            stack.extend(n.args)
        return sorted(set(out))

    def substitute(self, mapping: Dict[str, "ExprLike"]) -> "Expr":
        mp = {k: as_expr(v) for k, v in mapping.items()}
        memo: Dict[int, Expr] = {}
        for n in postorder(self):
            if n.kind == "sym":
                memo[id(n)] = mp.get(n.name, n)
            elif n.kind == "num":
                memo[id(n)] = n
            else:
                memo[id(n)] = rebuild(n, [memo[id(a)] for a in n.args])
        return memo[id(self)]

    def __repr__(self) -> str:
        return to_string(self)


ExprLike = Union[Expr, int, float]


def as_expr(v: ExprLike) -> Expr:
    if isinstance(v, Expr):
        return v
    if isinstance(v, (int, float)):
        return Expr.number(v)
    raise TypeError(f"cannot convert {type(v).__name__} to Expr")


def symb(name: str) -> Expr:
    return Expr.symbol(name)


def rebuild(n: Expr, args: Sequence[Expr]) -> Expr:
    if n.kind == "sum":
        return Expr.sum(args)
    if n.kind == "prod":
        return Expr.product(args)
    if n.kind == "div":
        return Expr.div(*args)
    if n.kind == "pow":
        return Expr.pow(*args)
    if n.kind == "call":
        return Expr.call(n.name, *args)
    return n


def postorder(root: Expr) -> List[Expr]:
    """Iterative post-order over the DAG, each node once (cf. codegen/traverse.rs:19-58)."""
    out: List[Expr] = []
    seen = set()
    stack: List[Tuple[Expr, bool]] = [(root, False)]
    while stack:
        n, done = stack.pop()
        if done:
            out.append(n)
            continue
        if id(n) in seen:
            continue
        seen.add(id(n))
        stack.append((n, True))
        for a in reversed(n.args):
            if id(a) not in seen:
                stack.append((a, False))
    return out


def to_string(e: Expr) -> str:
    memo: Dict[int, str] = {}
    for n in postorder(e):
        a = [memo[id(x)] for x in n.args]
        if n.kind == "num":
            v = n.value
            s = repr(int(v)) if v == int(v) and abs(v) < 1e15 else repr(v)
            memo[id(n)] = s if v >= 0 else f"({s})"
        elif n.kind == "sym":
            memo[id(n)] = n.name
        elif n.kind == "call":
            memo[id(n)] = f"{n.name}({', '.join(a)})"
        elif n.kind == "sum":
            memo[id(n)] = "(" + " + ".join(a) + ")"
        elif n.kind == "prod":
            memo[id(n)] = "(" + "*".join(a) + ")"
        elif n.kind == "div":
            memo[id(n)] = f"({a[0]}/{a[1]})"
        else:
            memo[id(n)] = f"({a[0]}^{a[1]})"
    return memo[id(e)]


# ------------------------------------------------------------------------------------------
# Parser (Pratt).  Grammar: + - * / ^ (right assoc, also **), unary -, calls f(a, b), numbers,
# identifiers, implicit multiplication between a number and an identifier/paren ("2x", "3(x+1)").
# ------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)|([A-Za-z_][A-Za-z_0-9]*)|(\*\*|[-+*/^(),]))")


class ParseError(ValueError):
    pass


def _tokenize(src: str) -> List[Tuple[str, str]]:
    toks: List[Tuple[str, str]] = []
    pos = 0
    src = src.rstrip()
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise ParseError(f"unexpected character {src[pos]!r} at {pos}")
        if m.group(1) is not None:
            toks.append(("num", m.group(1)))
        elif m.group(2) is not None:
            toks.append(("id", m.group(2)))
        else:
            toks.append(("op", "^" if m.group(3) == "**" else m.group(3)))
        pos = m.end()
    toks.append(("end", ""))
    return toks


class _Parser:
    def __init__(self, src: str):
        self.toks = _tokenize(src)
        self.i = 0

    def peek(self):
        return self.toks[self.i]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def expect(self, val: str):
        t = self.next()
        if t[1] != val:
            raise ParseError(f"expected {val!r}, got {t[1]!r}")

    def parse(self) -> Expr:
        e = self.expr(0)
        if self.peek()[0] != "end":
            raise ParseError(f"unexpected token {self.peek()[1]!r}")
        return e

    # binding powers: + - 10, * / 20, unary - 25, ^ 30 (right assoc)
    def expr(self, min_bp: int) -> Expr:
        kind, val = self.next()
        if kind == "num":
            lhs = Expr.number(float(val))
        elif kind == "id":
            if self.peek() == ("op", "("):
                self.next()
                args: List[Expr] = []
                if self.peek() != ("op", ")"):
                    while True:
                        args.append(self.expr(0))
                        if self.peek() == ("op", ","):
                            self.next()
                            continue
                        break
                self.expect(")")
                lhs = Expr.call(val, *args)
            else:
                lhs = Expr.symbol(val)
        elif (kind, val) == ("op", "("):
            lhs = self.expr(0)
            self.expect(")")
        elif (kind, val) == ("op", "-"):
            rhs = self.expr(25)
            lhs = Expr.product([-1.0, rhs])
        elif (kind, val) == ("op", "+"):
            lhs = self.expr(25)
        else:
            raise ParseError(f"unexpected token {val!r}")

        while True:
            kind, val = self.peek()
            if kind == "end" or (kind == "op" and val in "),"):
                break
            if kind == "op":
                if val in "+-":
                    bp = 10
                    if bp < min_bp:
                        break
                    self.next()
                    rhs = self.expr(bp + 1)
                    lhs = lhs + rhs if val == "+" else lhs - rhs
                elif val in "*/":
                    bp = 20
                    if bp < min_bp:
                        break
                    self.next()
                    rhs = self.expr(bp + 1)
                    lhs = lhs * rhs if val == "*" else lhs / rhs
                elif val == "^":
                    bp = 30
                    if bp < min_bp:
                        break
                    self.next()
                    rhs = self.expr(bp)  # right associative
                    lhs = Expr.pow(lhs, rhs)
                elif val == "(":
                    bp = 20  # implicit multiplication: 3(x+1)
                    if bp < min_bp:
                        break
                    rhs = self.expr(bp + 1)
                    lhs = lhs * rhs
                else:
                    raise ParseError(f"unexpected operator {val!r}")
            else:  # implicit multiplication: 2x, 2 sin(x)
                bp = 20
                if bp < min_bp:
                    break
                rhs = self.expr(bp + 1)
                lhs = lhs * rhs
        return lhs


def parse(src: str) -> Expr:
    """Parse an expression string (subset of ``src/parser/logic/pratt.rs``)."""
    return _Parser(src).parse()


# ------------------------------------------------------------------------------------------
# Differentiation
# ------------------------------------------------------------------------------------------
def _d_call(name: str, u: Expr) -> Optional[Expr]:
    """f'(u) for unary functions."""
    one = Expr.number(1.0)
    c = Expr.call
    table = {
        "sin": lambda: c("cos", u),
        "cos": lambda: -c("sin", u),
        "tan": lambda: one / c("cos", u) ** 2,
        "cot": lambda: -one / c("sin", u) ** 2,
        "sec": lambda: c("sec", u) * c("tan", u),
        "csc": lambda: -c("csc", u) * c("cot", u),
        "asin": lambda: one / c("sqrt", one - u ** 2),
        "acos": lambda: -one / c("sqrt", one - u ** 2),
        "atan": lambda: one / (one + u ** 2),
        "acot": lambda: -one / (one + u ** 2),
        "sinh": lambda: c("cosh", u),
        "cosh": lambda: c("sinh", u),
        "tanh": lambda: one - c("tanh", u) ** 2,
        "asinh": lambda: one / c("sqrt", u ** 2 + one),
        "acosh": lambda: one / c("sqrt", u ** 2 - one),
        "atanh": lambda: one / (one - u ** 2),
        "exp": lambda: c("exp", u),
        "ln": lambda: one / u,
        "log": lambda: one / u,
        "log1p": lambda: one / (one + u),
        "expm1": lambda: c("exp", u),
        "sqrt": lambda: one / (2.0 * c("sqrt", u)),
        "cbrt": lambda: one / (3.0 * c("cbrt", u) ** 2),
        "abs": lambda: c("signum", u),
        "erf": lambda: (2.0 / math.sqrt(math.pi)) * c("exp", -(u ** 2)),
        "erfc": lambda: (-2.0 / math.sqrt(math.pi)) * c("exp", -(u ** 2)),
        "gamma": lambda: c("gamma", u) * c("digamma", u),
        "lgamma": lambda: c("digamma", u),
        "digamma": lambda: c("trigamma", u),
        "trigamma": lambda: c("tetragamma", u),
        "sinc": lambda: (c("cos", u) * u - c("sin", u)) / u ** 2,
    }
    f = table.get(name)
    return f() if f else None


def diff(e: Expr, var: str) -> Expr:
    """d e / d var, computed bottom-up over the DAG (shared sub-derivatives stay shared)."""
    zero, one = Expr.number(0.0), Expr.number(1.0)
    d: Dict[int, Expr] = {}
    for n in postorder(e):
        if n.kind == "num":
            d[id(n)] = zero
        elif n.kind == "sym":
            d[id(n)] = one if n.name == var else zero
        elif n.kind == "sum":
            d[id(n)] = Expr.sum(d[id(a)] for a in n.args)
        elif n.kind == "prod":
            terms = []
            for i, a in enumerate(n.args):
                da = d[id(a)]
                if da is zero:
                    continue
                terms.append(Expr.product([da] + [b for j, b in enumerate(n.args) if j != i]))
            d[id(n)] = Expr.sum(terms)
        elif n.kind == "div":
            a, b = n.args
            da, db = d[id(a)], d[id(b)]
            if db is zero:
                d[id(n)] = da / b
            else:
                d[id(n)] = (da * b - a * db) / b ** 2
        elif n.kind == "pow":
            a, b = n.args
            da, db = d[id(a)], d[id(b)]
            if db is zero:
                if b.kind == "num":
                    d[id(n)] = Expr.product([b, Expr.pow(a, Expr.number(b.value - 1.0)), da])
                else:
                    d[id(n)] = Expr.product([b, Expr.pow(a, b - one), da])
            elif da is zero:
                d[id(n)] = Expr.product([n, Expr.call("ln", a), db])
            else:
                d[id(n)] = n * (db * Expr.call("ln", a) + b * da / a)
        elif n.kind == "call":
            if all(d[id(a)] is zero for a in n.args):
                d[id(n)] = zero
                continue
            if len(n.args) == 1:
                fp = _d_call(n.name, n.args[0])
                if fp is None:
                    raise NotImplementedError(f"d/dx of {n.name} is not implemented")
                d[id(n)] = fp * d[id(n.args[0])]
            elif n.name == "atan2":
                y, x = n.args
                d[id(n)] = (x * d[id(y)] - y * d[id(x)]) / (x ** 2 + y ** 2)
            else:
                raise NotImplementedError(f"d/dx of {n.name}/{len(n.args)} is not implemented")
        else:
            raise AssertionError(n.kind)
    return d[id(e)]


def gradient(e: Expr, variables: Sequence[str]) -> List[Expr]:
    return [diff(e, v) for v in variables]
