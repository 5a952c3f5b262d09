"""The Python surface mirrors the reference's PyO3 API for this path (symbanafis_b200/evaluator.py):
same names, argument meaning, return conventions and exception types as src/bindings/python/evaluator.rs:15-139,
parallel.rs:34-256 (stubs python/symb_anafis/__init__.pyi:454-545, 1100-1135).

The CPU half checks everything that happens before the device is touched (compilation, metadata, argument
validation); the GPU half replays the reference's own Python tests (tests/python/test_parallel_eval.py,
test_numpy_returns.py, test_compiled_evaluator.py) against the CUDA path."""
import math

import numpy as np
import pytest

from symbanafis_b200._lib import kernel_launch_count
from symbanafis_b200.evaluator import CompiledEvaluator, eval_f64, evaluate_parallel
from symbanafis_b200.expr import parse, symb


# ---- no device needed ------------------------------------------------------------------------------------------

def test_compiled_evaluator_metadata():
    ev = CompiledEvaluator(parse("x^2 + sin(x)*exp(-x)"), ["x"])
    assert ev.param_names() == ["x"] and ev.param_count() == 1
    # the golden lowering of config 1 (SURVEY section 8a note): Square, Sin, Builtin1(ExpNeg), MulAdd
    assert ev.instruction_count() == 4
    text = ev.disassemble()
    for op in ("Square", "Sin", "MulAdd"):
        assert op in text
    assert ev.workspace_size() >= 4
    # ... and the exact word stream SURVEY.md section 8a derives from the reference's lowering rules
    # (pow.rs:205-209 Square, func.rs:34-43 Builtin1{ExpNeg}, sum.rs:139-148 MulAdd; temporaries t0..t2 = R1..R3, result R4)
    assert [int(v) for v in ev.program.flat] == [20, 1, 0, 30, 2, 0, 38, 3, 26, 0, 15, 4, 2, 3, 1, 0]
    assert list(ev.program.consts) == [] and ev.program.param_count == 1 and list(ev.program.result_regs) == [4]


def test_auto_parameter_order_is_alphabetical_and_excludes_known_constants():
    ev = CompiledEvaluator(parse("b*a + pi + e"))  # compile_auto, api.rs:461-473
    assert ev.param_names() == ["a", "b"]


def test_compile_errors_are_runtime_errors():
    with pytest.raises(RuntimeError):
        CompiledEvaluator(parse("x + y"), ["x"])        # unbound variable
    with pytest.raises(RuntimeError):
        CompiledEvaluator(parse("nosuchfunction(x)"), ["x"])
    with pytest.raises(RuntimeError):
        CompiledEvaluator(parse("x"), ["x"], context=object())


def test_argument_validation_happens_before_any_device_work():
    with pytest.raises(ValueError):
        eval_f64(["x"], [["x"], ["y"]], [[[1.0]]])
    with pytest.raises(ValueError):
        evaluate_parallel(["x*y"], [["x", "y"]], [[[1.0]]])     # EvalColumnMismatch (parallel.rs:387-392)
    with pytest.raises(TypeError):
        CompiledEvaluator(parse("x"), ["x"]).eval_batch(["not numbers"])
    with pytest.raises(ValueError):
        CompiledEvaluator(parse("x"), ["x"]).eval_batch([np.arange(10.0)[::2]])  # not C-contiguous
    assert eval_f64([], [], []) == []                       # test_numpy_returns.py:43-47
    with pytest.raises(TypeError):
        evaluate_parallel([3.5], [["x"]], [[[1.0]]])           # "expressions must be str or Expr" (parallel.rs:52-55)
    with pytest.raises(ValueError):
        evaluate_parallel(["x*y"], [["x", "y"]], [[[1.0, 2.0], []]])  # empty column (parallel.rs:421-423)


def test_evaluate_parallel_all_skipped_points_take_the_symbolic_path_only():
    # tests/python/test_parallel_eval.py:159-187: a point with a None (Value::Skip) stays symbolic -- string in, string
    # out; Expr in, Expr out (parallel.rs:630-677); no numeric point -> no device work at all
    r = evaluate_parallel(["x + y"], [["x", "y"]], [[[1.0], [None]]])[0][0]
    assert isinstance(r, str) and "y" in r and "x" not in r
    x, y = symb("x"), symb("y")
    r = evaluate_parallel([x + y], [["x", "y"]], [[[1.0], [None]]])[0][0]
    from symbanafis_b200.expr import Expr
    assert isinstance(r, Expr) and r.variables() == ["y"]
    assert evaluate_parallel(["x"], [["x"]], [[[]]]) == [[]]
    # a custom hand-off (what a maintainer points at the reference's evaluate_slow_point)
    seen = []
    r = evaluate_parallel(["x*y"], [["x", "y"]], [[[None, None], [3.0]]],
                          symbolic_fallback=lambda e, names, point, was_string: seen.append((names, point)) or "handled")
    assert r == [["handled", "handled"]] and seen == [(["x", "y"], [None, 3.0])] * 2


# ---- the reference's Python tests against the CUDA path --------------------------------------------------------
gpu = pytest.mark.gpu


@gpu
def test_evaluate_parallel_basic_cases():
    # tests/python/test_parallel_eval.py:19-57
    assert evaluate_parallel(["x^2"], [["x"]], [[[1.0, 2.0, 3.0]]]) == [[1.0, 4.0, 9.0]]
    assert evaluate_parallel(["x + y"], [["x", "y"]], [[[1.0, 2.0], [10.0, 20.0]]]) == [[11.0, 22.0]]
    r = evaluate_parallel(["x^2", "x + 1"], [["x"], ["x"]], [[[1.0, 2.0, 3.0]], [[1.0, 2.0, 3.0]]])
    assert r == [[1.0, 4.0, 9.0], [2.0, 3.0, 4.0]]


@gpu
def test_evaluate_parallel_numpy_and_expr_inputs():
    # tests/python/test_parallel_eval.py:62-128
    assert evaluate_parallel(["x^2"], [["x"]], [[np.array([1.0, 2.0, 3.0, 4.0])]])[0] == [1.0, 4.0, 9.0, 16.0]
    x, y = np.array([1.0, 2.0, 3.0]), np.array([10.0, 20.0, 30.0])
    assert evaluate_parallel(["x * y"], [["x", "y"]], [[x, y]])[0] == [10.0, 40.0, 90.0]
    assert evaluate_parallel(["x + y"], [["x", "y"]], [[x, [10.0, 20.0, 30.0]]])[0] == [11.0, 22.0, 33.0]
    e = symb("x") ** 2
    assert evaluate_parallel([e], [["x"]], [[[1.0, 2.0, 3.0]]])[0] == [1.0, 4.0, 9.0]
    r = evaluate_parallel([e, "x + 1"], [["x"], ["x"]], [[[1.0, 2.0]], [[1.0, 2.0]]])
    assert r == [[1.0, 4.0], [2.0, 3.0]]
    # a shorter value column is broadcast from its last value (parallel.rs:440-465)
    assert evaluate_parallel(["x + y"], [["x", "y"]], [[[1.0, 2.0, 3.0], [10.0]]])[0] == [11.0, 12.0, 13.0]


@gpu
def test_evaluate_parallel_mixed_points_numeric_on_the_gpu_symbolic_handed_off():
    # tests/python/test_parallel_eval.py:131-144 (parallel.rs:489-613): numeric points through the compiled evaluator,
    # here gathered into ONE launch; points with a skipped value through the symbolic slow path
    before = kernel_launch_count()
    r = evaluate_parallel(["x * y"], [["x", "y"]], [[[2.0, None, 4.0, None, 1.5], [3.0, 5.0]]])
    assert kernel_launch_count() - before == 1
    assert r[0][0] == 6.0 and r[0][2] == 20.0 and r[0][4] == 7.5 and all(isinstance(r[0][i], float) for i in (0, 2, 4))
    assert isinstance(r[0][1], str) and "x" in r[0][1] and "5" in r[0][1] and isinstance(r[0][3], str)
    # zero-variable expression: one value (parallel.rs:399-409)
    assert evaluate_parallel(["2 + 3"], [[]], [[]]) == [[5.0]]


@gpu
def test_eval_f64_return_conventions():
    # tests/python/test_numpy_returns.py:6-75
    x = np.array([1.0, 2.0, 3.0])
    y = np.array([10.0, 20.0, 30.0])
    res = eval_f64([parse("x^2"), parse("x + y")], [["x"], ["x", "y"]], [[x], [x, y]])
    assert isinstance(res, np.ndarray) and res.shape == (2, 3) and res.dtype == np.float64
    np.testing.assert_allclose(res[0], x ** 2)
    np.testing.assert_allclose(res[1], x + y)
    empty = eval_f64([parse("x")], [["x"]], [[np.array([], dtype=np.float64)]])
    assert isinstance(empty, np.ndarray) and empty.shape == (1, 0)
    lists = eval_f64([parse("x * 2")], [["x"]], [[[1.0, 2.0, 3.0]]])
    assert isinstance(lists, list) and lists == [[2.0, 4.0, 6.0]]
    with pytest.raises(ValueError):
        eval_f64([parse("x + y")], [["x", "y"]], [[x, y[:2]]])   # EvalColumnLengthMismatch -> ValueError


@gpu
def test_eval_f64_fuses_expressions_over_the_same_columns_into_one_pass():
    # the reference evaluates [dx, dy, dz] as three independent passes (api.rs:365-375, aizawa_flow.py:103-112)
    rng = np.random.default_rng(4)
    x, y, z = (rng.normal(size=50_000) for _ in range(3))
    exprs = ["(z-0.7)*x-3.5*y", "3.5*x+(z-0.7)*y", "0.6+0.95*z-z^3/3-(x^2+y^2)*(1+0.25*z)+0.1*z*x^3"]
    names = [["x", "y", "z"]] * 3
    eval_f64(exprs, names, [[x, y, z]] * 3)           # compile + cache
    before = kernel_launch_count()
    res = eval_f64(exprs, names, [[x, y, z]] * 3)
    assert kernel_launch_count() - before == 1        # ONE launch for the three outputs
    np.testing.assert_allclose(res[0], (z - 0.7) * x - 3.5 * y, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(res[1], 3.5 * x + (z - 0.7) * y, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(res[2], 0.6 + 0.95 * z - z ** 3 / 3 - (x ** 2 + y ** 2) * (1 + 0.25 * z) + 0.1 * z * x ** 3,
                               rtol=1e-11, atol=1e-13)


@gpu
def test_compiled_evaluator_batch_and_scalar_conventions():
    # tests/python/test_compiled_evaluator.py:27-110
    ev = CompiledEvaluator(parse("x^2 + 1"), ["x"])
    assert ev.evaluate([2.0]) == 5.0 and isinstance(ev.evaluate(np.array([2.0])), float)
    assert ev.eval_batch([[1.0, 2.0, 3.0, 4.0]]) == [2.0, 5.0, 10.0, 17.0]            # list in -> list out
    out = ev.eval_batch([np.array([1.0, 2.0])])
    assert isinstance(out, np.ndarray) and out.tolist() == [2.0, 5.0]                   # ndarray in -> ndarray out
    big = CompiledEvaluator(parse("sin(x) + cos(x)"), ["x"]).eval_batch([np.linspace(0, 2 * math.pi, 10000)])
    assert len(big) == 10000 and abs(big[0] - 1.0) < 1e-10
    assert CompiledEvaluator(parse("42"), []).evaluate([]) == 42.0
    assert abs(CompiledEvaluator(parse("sin(pi/2)"), []).evaluate([]) - 1.0) < 1e-10


def test_package_exports_the_reference_names_lazily():
    import importlib
    import symbanafis_b200
    importlib.reload(symbanafis_b200)
    assert {"CompiledEvaluator", "eval_f64", "evaluate_parallel"} <= set(dir(symbanafis_b200))
    from symbanafis_b200 import CompiledEvaluator as A
    from symbanafis_b200.evaluator import CompiledEvaluator as B
    assert A is B
    import pytest as _pytest
    with _pytest.raises(AttributeError):
        symbanafis_b200.no_such_name
