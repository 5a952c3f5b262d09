"""Pins the oracle -- and, on the GPU, the CUDA path -- on the reference's own known answers.

tests/golden/reference_kats.json is produced by tests/golden/make_golden.py from the reference's tests
(src/tests/precision_audit.rs, drivers/tests.rs, tests/python/*.py; file:line kept per case).  The
reference asserts those answers on constant expressions; here every numeric literal that is a function
argument is lifted into a PARAMETER so that the value really goes through the bytecode interpreter (the
oracle on CPU, the sm_100a kernel on the GPU) instead of being folded by the host compiler.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import OracleProgram
from symbanafis_b200.compiler import compile_exprs
from symbanafis_b200.expr import Expr, parse, postorder, rebuild

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_kats.json")) as _f:
    GOLDEN = json.load(_f)


def lift_literals(expr: Expr):
    """Replaces numeric literals that are direct arguments of a function call by parameters q0, q1, ..."""
    memo, values = {}, []
    for n in postorder(expr):
        if n.kind == "call":
            args = []
            for a in n.args:
                a2 = memo[id(a)]
                if a2.kind == "num":
                    values.append(a2.value)
                    a2 = Expr.symbol(f"q{len(values) - 1}")
                args.append(a2)
            memo[id(n)] = Expr.call(n.name, *args)
        elif n.kind in ("num", "sym"):
            memo[id(n)] = n
        else:
            memo[id(n)] = rebuild(n, [memo[id(a)] for a in n.args])
    return memo[id(expr)], values


def audit_ok(computed: float, expected: float, rel_tol: float, abs_tol: float) -> bool:
    """check_rel_error, src/tests/precision_audit.rs:32-57."""
    if abs(expected) < abs_tol:
        return abs(computed - expected) < abs_tol
    return abs((computed - expected) / expected) < rel_tol


def approx_eq(a: float, b: float, eps: float) -> bool:
    """tests/python/test_compiled_evaluator.py:16-22; eps 0 = exact."""
    if math.isnan(a) and math.isnan(b):
        return True
    if math.isinf(a) and math.isinf(b):
        return (a > 0) == (b > 0)
    if eps == 0.0:
        return a == b
    return abs(a - b) < eps * max(abs(a), abs(b), 1.0)


def _audit_program(case):
    e, values = lift_literals(parse(case["expr"].replace("_", "")))
    names = [f"q{i}" for i in range(len(values))]
    return compile_exprs([e], names), values


AUDIT = GOLDEN["precision_audit"]
TRANS = GOLDEN["transcribed"]


def test_fixture_is_complete():
    assert len(AUDIT) >= 80 and len(TRANS) >= 30
    assert all(c["source"].startswith("src/tests/precision_audit.rs:") for c in AUDIT)


@pytest.mark.parametrize("case", AUDIT, ids=[c["expr"] for c in AUDIT])
def test_oracle_reproduces_precision_audit(case):
    prog, values = _audit_program(case)
    got = float(OracleProgram.from_program(prog).evaluate(values)[0])
    assert audit_ok(got, case["expected"], case["rel_tol"], case["abs_tol_near_zero"]), (case, got)


@pytest.mark.parametrize("case", TRANS, ids=[f'{c["expr"]}@{i}' for i, c in enumerate(TRANS)])
def test_oracle_reproduces_transcribed_known_answers(case):
    prog = compile_exprs([parse(case["expr"])], case["params"])
    o = OracleProgram.from_program(prog)
    if case["kind"] == "evaluate":
        got = [float(o.evaluate(case["inputs"])[0])]
        exp = [case["expected"]]
    else:
        cols = [np.asarray(c, dtype=np.float64) for c in case["columns"]]
        got = list(o.eval_batch(cols, len(case["expected"]))[0])
        exp = case["expected"]
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        assert approx_eq(float(g), e, case["tol"]), (case, got)


# ---- the same fixtures through the CUDA path (C ABI -> sm_100a interpreter) -----------------------------

@pytest.mark.gpu
def test_device_reproduces_precision_audit():
    from symbanafis_b200.evaluator import CompiledEvaluator
    bad = []
    for case in AUDIT:
        prog, values = _audit_program(case)
        ev = CompiledEvaluator.from_program(prog)
        got = ev.evaluate(values)
        if not audit_ok(got, case["expected"], case["rel_tol"], case["abs_tol_near_zero"]):
            bad.append((case["expr"], got, case["expected"]))
        # a batch of copies of the same point must agree with the single-point call bit for bit
        if values:
            cols = [np.full(300, v) for v in values]
            many = np.asarray(ev.eval_batch(cols))
            assert many.shape == (300,) and (many.view(np.uint64) == np.float64(got).view(np.uint64)).all() \
                or (np.isnan(got) and np.isnan(many).all()), case["expr"]
    assert not bad, bad


@pytest.mark.gpu
def test_device_reproduces_transcribed_known_answers():
    from symbanafis_b200.evaluator import CompiledEvaluator
    for case in TRANS:
        ev = CompiledEvaluator(case["expr"], case["params"])
        if case["kind"] == "evaluate":
            got, exp = [ev.evaluate(case["inputs"])], [case["expected"]]
        else:
            got = list(ev.eval_batch([list(c) for c in case["columns"]]))
            exp = case["expected"]
            as_np = ev.eval_batch([np.asarray(c, dtype=np.float64) for c in case["columns"]])
            assert isinstance(as_np, np.ndarray) and list(as_np) == got  # ndarray in -> ndarray out
        assert len(got) == len(exp), case
        for g, e in zip(got, exp):
            assert approx_eq(float(g), e, case["tol"]), (case, got)
