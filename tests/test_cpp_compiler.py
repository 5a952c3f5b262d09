"""The C++ expression -> bytecode compiler behind the C ABI (csrc/compile.cpp, saf_compile_exprs) against the Python
harness compiler (symbanafis_b200/compiler.py): the same parser, the same lowering rules, the same value numbering,
register allocation and encoding -- the payloads must be identical word for word (constants bit for bit), over the
source expressions of the five workloads, the re-parsed terrain gradient (2666 instructions), the reference's
SIMD-surface corpus (src/tests/fuzz_evaluator.rs:254-291), parser conventions and random expressions; and the
C1 stream is the one SURVEY.md section 8a derives from the reference's lowering rules."""
import json
import os

import numpy as np
import pytest

from oracle import OracleProgram
from symbanafis_b200 import workloads
from symbanafis_b200._lib import DeviceProgramHandle, SafError
from symbanafis_b200.compiler import compile_exprs
from symbanafis_b200.expr import diff, parse, to_string
from test_gpu_parity_ranges import SIMD_SURFACE
from util import bits_equal


def both(exprs, params):
    h = DeviceProgramHandle.compile(exprs, params)
    flat, consts, pool, pc, ws, rr = h.reference_payload()
    py = compile_exprs([parse(e) for e in exprs], params)
    return (flat, consts, pool, pc, ws, rr), py, h


def assert_same(exprs, params):
    (flat, consts, pool, pc, ws, rr), py, _ = both(exprs, params)
    assert list(flat) == list(py.flat), f"bytecode differs for {exprs}"
    assert consts.tobytes() == py.consts.tobytes(), f"constants differ for {exprs}"
    assert list(pool) == list(py.arg_pool) and pc == py.param_count and ws == py.workspace_size and rr == py.result_regs


def test_config_1_word_stream():
    (flat, consts, pool, pc, ws, rr), _, _ = both(["x^2 + sin(x)*exp(-x)"], ["x"])
    assert list(flat) == [20, 1, 0, 30, 2, 0, 38, 3, 26, 0, 15, 4, 2, 3, 1, 0] and len(consts) == 0 and rr == [4]


def test_workload_source_expressions():
    assert_same(["sin(-1.4*y) + 1.0*cos(-1.4*x)", "sin(1.6*x) + 0.7*cos(1.6*y)"], ["x", "y"])
    assert_same(["(z - 0.7) * x - 3.5 * y", "3.5 * x + (z - 0.7) * y",
                 "0.6 + 0.95*z - z^3/3 - (x^2 + y^2) * (1 + 0.25*z) + 0.1 * z * x^3"], ["x", "y", "z"])
    with open(os.path.join(os.path.dirname(workloads.__file__), "data", "double_pendulum_rhs.json")) as f:
        rhs = json.load(f)
    assert_same([rhs["a1"], rhs["a2"]], rhs["params"])
    assert_same([workloads.TERRAIN_BASE], ["x", "y"])


def test_terrain_gradient_reparsed_is_identical_and_evaluates_like_the_workload():
    h = workloads.terrain_expr(110)
    gx, gy = to_string(diff(h, "x")), to_string(diff(h, "y"))
    (flat, consts, pool, pc, ws, rr), py, _ = both([gx, gy], ["x", "y"])
    assert len(flat) > 8000  # thousands of instructions, AddN pools
    assert list(flat) == list(py.flat) and consts.tobytes() == py.consts.tobytes() and list(pool) == list(py.arg_pool)
    assert ws == py.workspace_size and rr == py.result_regs
    # the printed-and-reparsed gradient computes what the workload's program computes (oracle on both)
    w = workloads.get("terrain_gradient")
    cols = w.make_inputs(257, 9)
    a = OracleProgram(flat, consts, pool, pc, ws, rr).eval_batch(cols, 257)
    b = OracleProgram.from_program(w.program).eval_batch(cols, 257)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("expr,names", SIMD_SURFACE, ids=[e for e, _ in SIMD_SURFACE])
def test_reference_corpus(expr, names):
    assert_same([expr], names)


PARSER_CASES = [
    ("2x + 3(x+1)", ["x"]), ("2 sin(x) cos(x)", ["x"]), ("-x^2", ["x"]), ("x**3 - -x", ["x"]), ("2^3^2 * x", ["x"]),
    ("1/(exp(x)-1)", ["x"]), ("sin(x)/x", ["x"]), ("x/4 + y/0.5", ["x", "y"]), ("e^x + pi*x", ["x"]), ("exp(-x^2) + exp(x^2)", ["x"]),
    ("exp(-2*x*y)", ["x", "y"]), ("exp(-x/y)", ["x", "y"]), ("x^0.5 + x^-0.5 + x^1.5 + x^-1.5 + x^7 + x^-2 + x^-3 + x^2.5", ["x"]),
    ("log2(x) + log10(x) + log(x)", ["x"]), ("a*b - c", ["a", "b", "c"]), ("c - a*b", ["a", "b", "c"]), ("a*b + c*d + e*f - g", list("abcdefg")),
    ("-(a*b) - c - d", ["a", "b", "c", "d"]), ("a+b+c+d+e+f+g", list("abcdefg")), ("a*b*c*d*e*f", list("abcdef")),
    ("sin(x)^2 + cos(x)^2", ["x"]), ("3 + 4*2", []), ("x*0 + 1.5e-3", ["x"]), (".5x + 1.e2", ["x"]), ("pi", ["pi"]),
    ("abs(x) + sign(x) + sgn(x)", ["x"]), ("sqrt(2) * x", ["x"]), ("x - x", ["x"]), ("(x+1)*(x+1)", ["x"]),
]


@pytest.mark.parametrize("expr,names", PARSER_CASES, ids=[e for e, _ in PARSER_CASES])
def test_parser_and_lowering_conventions(expr, names):
    assert_same([expr], names)


def _random_expr(rng, depth, names):
    if depth == 0 or rng.random() < 0.15:
        r = rng.random()
        if r < 0.5:
            return str(rng.choice(names))
        if r < 0.8:
            return repr(round(float(rng.uniform(-3, 3)), 3))
        return str(int(rng.integers(0, 5)))
    k = rng.integers(0, 8)
    a = _random_expr(rng, depth - 1, names)
    b = _random_expr(rng, depth - 1, names)
    if k == 0: return f"({a} + {b})"
    if k == 1: return f"({a} - {b})"
    if k == 2: return f"({a} * {b})"
    if k == 3: return f"({a} / {b})"
    if k == 4: return f"({a})^{int(rng.integers(-3, 5))}"
    if k == 5: return f"{rng.choice(['sin', 'cos', 'exp', 'ln', 'sqrt', 'tanh', 'atan', 'erf', 'abs'])}({a})"
    if k == 6: return f"-({a})"
    return f"({a} + {b} + {_random_expr(rng, depth - 1, names)} - {_random_expr(rng, depth - 1, names)})"


@pytest.mark.parametrize("seed", range(60))
def test_random_expressions(seed):
    rng = np.random.default_rng(4200 + seed)
    names = ["x", "y", "z"][: int(rng.integers(1, 4))]
    exprs = [_random_expr(rng, int(rng.integers(2, 6)), names) for _ in range(int(rng.integers(1, 4)))]
    assert_same(exprs, names)


def test_compiled_program_runs_in_the_oracle_like_the_python_one():
    (flat, consts, pool, pc, ws, rr), py, _ = both(["x^2 + sin(x)*exp(-x)", "x*y - 1/(1+y^2)"], ["x", "y"])
    rng = np.random.default_rng(3)
    cols = [rng.uniform(-2, 2, 300), rng.uniform(-2, 2, 300)]
    a = OracleProgram(flat, consts, pool, pc, ws, rr).eval_batch(cols, 300)
    b = OracleProgram.from_program(py).eval_batch(cols, 300)
    for x, y in zip(a, b):
        assert bits_equal(x, y)


def test_errors():
    for bad, params in (("x + y", ["x"]), ("nosuchfunction(x)", ["x"]), ("sin(x, x)", ["x"])):
        with pytest.raises(SafError) as e:
            DeviceProgramHandle.compile([bad], params)
        assert "Unbound" in e.value.message or "Unsupported" in e.value.message
    for bad in ("x +", "(x", "x $ y", ")"):
        with pytest.raises(SafError) as e:
            DeviceProgramHandle.compile([bad], ["x", "y"])
        assert "parse error" in e.value.message
