"""The five north-star workloads (BASELINE.json ``configs``): programs + synthetic inputs.

Definitions follow the reference's example scripts:

* C1 ``x^2 + sin(x)*exp(-x)``, 1M points                    -- eval_parallel! plumbing case
* C2 Clifford map step, 1M particles x 1000 iterations      -- examples/python/clifford_benchmark.py:25,107-108,133-134
* C3 Aizawa field + Euler step, 500k particles              -- examples/python/aizawa_flow.py:24,28,99-101,121-123,163-165
* C4 double pendulum, damped RK4 step, 50k pendulums        -- examples/python/double_pendulum_benchmark.py:25,32-35,317-365,377-381
* C5 gradient of a large synthetic terrain, 100M points     -- examples/python/gradient_descent.py:33,192-201 scaled to
                                                               ``dump_large_expr`` size (the reference's big_expr.txt is absent)

Each workload returns a :class:`Workload` with the fused multi-output program (reference ISA), the
separately compiled per-expression programs the reference would run, and an input generator.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .compiler import compile_exprs
from .expr import Expr, diff, parse, symb
from .fuse import euler_step, merge_programs, rk4_pendulum_step
from .isa import Program

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


@dataclass
class Workload:
    name: str
    description: str
    params: List[str]
    program: Program                      # fused program: one pass = one unit of the metric
    parts: List[Program]                  # what the reference evaluates separately (one per expression)
    n_points: int                         # BASELINE.json size
    iterations: int                       # map iterations / integration steps of the config (1 = single pass)
    make_inputs: Callable[[int, int], List[np.ndarray]]  # (n, seed) -> columns
    iterated: bool = False                # True when results feed back into the parameters
    notes: str = ""
    extra: Dict = field(default_factory=dict)


# ---- C1 ---------------------------------------------------------------------------------------------
def single_expression() -> Workload:
    e = parse("x^2 + sin(x)*exp(-x)")
    prog = compile_exprs([e], ["x"])

    def inputs(n: int, seed: int = 1):
        return [np.random.default_rng(seed).uniform(-2.0, 2.0, n)]

    return Workload("single_expr", "x^2 + sin(x)*exp(-x) over 1M points", ["x"], prog, [prog], 1_000_000, 1, inputs)


# ---- C2 ---------------------------------------------------------------------------------------------
CLIFFORD = dict(a=-1.4, b=1.6, c=1.0, d=0.7)  # clifford_benchmark.py:25


def clifford() -> Workload:
    a, b, c, d = (CLIFFORD[k] for k in "abcd")
    ex = parse(f"sin({a}*y) + {c}*cos({a}*x)")
    ey = parse(f"sin({b}*x) + {d}*cos({b}*y)")
    names = ["x", "y"]
    parts = [compile_exprs([ex], names), compile_exprs([ey], names)]
    prog = compile_exprs([ex, ey], names)

    def inputs(n: int, seed: int = 2):
        rng = np.random.default_rng(seed)
        return [rng.uniform(-2.0, 2.0, n), rng.uniform(-2.0, 2.0, n)]

    return Workload("clifford", "Clifford attractor map step (2 outputs), 1M particles x 1000 iterations",
                    names, prog, parts, 1_000_000, 1000, inputs, iterated=True)


# ---- C3 ---------------------------------------------------------------------------------------------
AIZAWA = dict(A=0.95, B=0.7, C=0.6, D=3.5, E=0.25, F=0.1)  # aizawa_flow.py:28
AIZAWA_DT = 0.01                                           # aizawa_flow.py:24


def aizawa() -> Workload:
    A, B, C, D, E, F = (AIZAWA[k] for k in "ABCDEF")
    names = ["x", "y", "z"]
    edx = parse(f"(z - {B}) * x - {D} * y")
    edy = parse(f"{D} * x + (z - {B}) * y")
    edz = parse(f"{C} + {A}*z - z^3/3 - (x^2 + y^2) * (1 + {E}*z) + {F} * z * x^3")
    parts = [compile_exprs([e], names) for e in (edx, edy, edz)]
    fld = compile_exprs([edx, edy, edz], names)
    prog = euler_step(fld, AIZAWA_DT, names)

    def inputs(n: int, seed: int = 3):
        rng = np.random.default_rng(seed)
        return [rng.normal(0, 0.1, n), rng.normal(0, 0.1, n), rng.normal(0, 0.5, n)]

    return Workload("aizawa", "Aizawa attractor Euler step (3-output field), 500k particles", names, prog, parts,
                    500_000, 600, inputs, iterated=True, extra={"field": fld, "dt": AIZAWA_DT})


# ---- C4 ---------------------------------------------------------------------------------------------
PENDULUM_DT, PENDULUM_DAMPING = 0.01, 0.9995  # double_pendulum_benchmark.py:25,35


def double_pendulum() -> Workload:
    with open(os.path.join(_DATA, "double_pendulum_rhs.json")) as f:
        rhs = json.load(f)
    names = rhs["params"]
    a1, a2 = parse(rhs["a1"]), parse(rhs["a2"])
    parts = [compile_exprs([a1], names), compile_exprs([a2], names)]
    accel = compile_exprs([a1, a2], names)
    prog = rk4_pendulum_step(accel, PENDULUM_DT, PENDULUM_DAMPING)

    def inputs(n: int, seed: int = 42):
        rng = np.random.RandomState(seed)  # the script uses np.random.seed(42) (:377-381)
        th1 = rng.uniform(2.5, 3.0, n)
        th2 = rng.uniform(0.5, 1.5, n)
        return [th1, th2, np.zeros(n), np.zeros(n)]

    return Workload("double_pendulum", "Double pendulum RK4, symbolically derived EOM, 50k pendulums, fused 4-stage step",
                    names, prog, parts, 50_000, 600, inputs, iterated=True,
                    extra={"accel": accel, "dt": PENDULUM_DT, "damping": PENDULUM_DAMPING})


# ---- C5 ---------------------------------------------------------------------------------------------
TERRAIN_BASE = ("sin(x) * cos(y) + 0.1 * (x^2 + y^2) - 0.2 * cos(1.5 * x - 1.0) * sin(2.0 * y) "
                "+ 0.2 * erf(x/2.0) * erf(y/2.0)")  # gradient_descent.py:33


def _lcg(seed: int):
    state = seed & 0xFFFFFFFF
    while True:  # Numerical Recipes LCG, 32 bit
        state = (1664525 * state + 1013904223) & 0xFFFFFFFF
        yield state / 4294967296.0


def terrain_expr(n_terms: int, seed: int = 12345) -> Expr:
    """Base terrain + sum_k [A_k sin(f_k x+p_k) cos(g_k y+q_k) + B_k exp(-((x-c_k)^2+(y-d_k)^2)/w_k)]."""
    x, y = symb("x"), symb("y")
    u = _lcg(seed)
    e = parse(TERRAIN_BASE)
    terms = [e]

    def r(lo, hi):
        return round(lo + (hi - lo) * next(u), 6)

    for _ in range(n_terms):
        A, f, p, g, q = r(-0.3, 0.3), r(0.2, 3.0), r(-3.0, 3.0), r(0.2, 3.0), r(-3.0, 3.0)
        B, c, d, w = r(-0.5, 0.5), r(-6.0, 6.0), r(-6.0, 6.0), r(0.5, 4.0)
        wave = A * Expr.call("sin", f * x + p) * Expr.call("cos", g * y + q)
        bump = B * Expr.call("exp", -(((x - c) ** 2 + (y - d) ** 2) / w))
        terms += [wave, bump]
    return Expr.sum(terms)


def terrain_gradient(n_terms: int = 110) -> Workload:
    names = ["x", "y"]
    h = terrain_expr(n_terms)
    gx, gy = diff(h, "x"), diff(h, "y")
    parts = [compile_exprs([gx], names), compile_exprs([gy], names)]
    prog = compile_exprs([gx, gy], names)

    def inputs(n: int, seed: int = 42):
        rng = np.random.RandomState(seed)  # gradient_descent.py:192-201
        na = n // 2
        xa, ya = rng.normal(-2.5, 0.8, na), rng.normal(2.5, 0.8, na)
        nb = n - na
        xb, yb = rng.normal(2.5, 0.8, nb), rng.normal(-2.5, 0.8, nb)
        return [np.concatenate([xa, xb]), np.concatenate([ya, yb])]

    return Workload("terrain_gradient",
                    f"Fused (d/dx, d/dy) of a {n_terms}-term synthetic terrain (dump_large_expr scale), 100M points",
                    names, prog, parts, 100_000_000, 1, inputs, extra={"n_terms": n_terms})


WORKLOADS = {
    "single_expr": single_expression,
    "clifford": clifford,
    "aizawa": aizawa,
    "double_pendulum": double_pendulum,
    "terrain_gradient": terrain_gradient,
}


def get(name: str) -> Workload:
    return WORKLOADS[name]()
