"""Parity over the WHOLE range of the lean transcendental kernels and of the libm-backed builtins (round-1 verdict:
the 4-ulp bar was only exercised on |x| < 50 / 1e4 and the builtins were held to 2e-11 relative).

* Sin / Cos / SinCos: log-uniform |x| in [1e-300, 2^20) -- the whole fast path of csrc/leanmath.cuh -- plus 10 k points
  within 1e-9 (relative) of k pi/2 for |k| <= 3e5, where a three-term Cody-Waite reduction cancels the most, plus both
  sides of the 2^20 switch to the CUDA math library;
* every Builtin1 whose reference implementation is ONE libm call (builtins.rs:40-70), ExpNeg included: <= 4 ulp;
* the sign of zero: Sin / Exp / Sqrt / Neg / SinCos / the odd builtins at +-0 (util.ulp_distance cannot see it);
* the reference's SIMD-surface corpus (src/tests/fuzz_evaluator.rs:250-327): 36 expressions on x, y in [-4, 4).
"""
import numpy as np
import pytest

from oracle import OracleProgram
from symbanafis_b200.compiler import compile_exprs
from symbanafis_b200.expr import parse
from symbanafis_b200.isa import FNOP
from test_gpu_parity import gpu_eval, single_op
from util import assert_close, ulp_distance

pytestmark = pytest.mark.gpu

TWO20 = 1048576.0


def _trig_points():
    rng = np.random.default_rng(20)
    mag = np.exp(rng.uniform(np.log(1e-300), np.log(TWO20), 60000))
    x = np.where(rng.random(mag.size) < 0.5, -mag, mag)
    x = x[np.abs(x) < TWO20]
    k = rng.integers(-300000, 300001, 10000).astype(np.float64)
    near = k * (np.pi / 2) * (1.0 + rng.uniform(-1e-9, 1e-9, k.size))
    exact = np.round(np.arange(-40, 41) * (np.pi / 2), 15)
    edge = np.array([np.nextafter(TWO20, 0), TWO20, np.nextafter(TWO20, np.inf), -np.nextafter(TWO20, 0), -TWO20,
                     1.5 * TWO20, 1e15, 1e22, 1e300, 5e-324, -5e-324, 2.2250738585072014e-308])
    dense = rng.uniform(-TWO20, TWO20, 40000)
    return np.concatenate([x, near, exact, edge, dense])


@pytest.mark.parametrize("op", ["Sin", "Cos"])
def test_sin_cos_within_4_ulp_over_the_whole_fast_range(op):
    x = _trig_points()
    prog = single_op(op, 1, 0, n_in=1)
    ref = OracleProgram.from_program(prog).eval_batch([x], x.size)[0]
    got = gpu_eval(prog, [x], x.size)[0]
    d = ulp_distance(got, ref)
    assert d.max() <= 4, f"{op}: {d.max()} ulp at x={x[np.argmax(d)]!r}"
    assert np.array_equal(np.signbit(got), np.signbit(ref)), f"{op}: sign differs at {x[np.signbit(got) != np.signbit(ref)][:3]}"


def test_sincos_within_4_ulp_over_the_whole_fast_range():
    x = _trig_points()
    prog = single_op("SinCos", 1, 2, 0, n_in=1, n_out=2)
    ref = OracleProgram.from_program(prog).eval_batch([x], x.size)
    got = gpu_eval(prog, [x], x.size)
    for g, r, what in zip(got, ref, ("sin", "cos")):
        d = ulp_distance(g, r)
        assert d.max() <= 4, f"sincos.{what}: {d.max()} ulp at x={x[np.argmax(d)]!r}"
        assert np.array_equal(np.signbit(g), np.signbit(r)), f"sincos.{what}: sign differs"


@pytest.mark.parametrize("op", ["Exp", "ExpSqrNeg", "ExpSqr"])
def test_exp_family_within_4_ulp_including_the_range_switch(op):
    rng = np.random.default_rng(21)
    lim = 700.0 if op == "Exp" else 26.4
    x = np.concatenate([rng.uniform(-lim, lim, 40000), rng.uniform(-1e-3, 1e-3, 5000),
                        [np.nextafter(700.0, 0), 700.0, np.nextafter(700.0, np.inf), -700.0, -745.2, 709.78, 710.0, -800.0,
                         0.0, -0.0, 5e-324]]) if op == "Exp" else np.concatenate(
        [rng.uniform(-lim, lim, 40000), [26.45, 26.46, 26.5, 27.3, 0.0, -0.0]])
    prog = single_op(op, 1, 0, n_in=1)
    ref = OracleProgram.from_program(prog).eval_batch([x], x.size)[0]
    got = gpu_eval(prog, [x], x.size)[0]
    d = ulp_distance(got, ref)
    assert d.max() <= 4, f"{op}: {d.max()} ulp at x={x[np.argmax(d)]!r}"


# Builtin1 functions that are ONE libm call in the reference (builtins.rs:40-70): the per-transcendental bar applies.
# (Reciprocal forms cot/sec/csc/coth/sech/csch add one correctly rounded division: <= 4 ulp as well, away from the poles.)
LIBM_BUILTIN1 = {
    "tan": (-1.55, 1.55), "asin": (-1, 1), "acos": (-1, 1), "atan": (-1e6, 1e6), "sinh": (-30, 30), "cosh": (-30, 30),
    "tanh": (-20, 20), "expm1": (-30, 30), "exp_neg": (-700, 700), "log1p": (-0.99, 1e6), "cbrt": (-1e6, 1e6),
    "cot": (0.05, 3.09), "sec": (-1.5, 1.5), "csc": (0.05, 3.09), "coth": (0.01, 20), "sech": (-30, 30), "csch": (0.01, 30),
    "exp_polar": (-700, 700),
}  # not here: acot = pi/2 - atan x (builtins.rs:47) SUBTRACTS -- a 1-ulp difference in atan is amplified by (pi/2) / acot x


@pytest.mark.parametrize("fn", sorted(LIBM_BUILTIN1))
def test_libm_backed_builtin1_within_4_ulp(fn):
    rng = np.random.default_rng(sum(map(ord, fn)))
    lo, hi = LIBM_BUILTIN1[fn]
    x = np.concatenate([rng.uniform(lo, hi, 30000), rng.uniform(max(lo, -1e-3), min(hi, 1e-3), 3000) if lo < 0 < hi else [],
                        [0.0, -0.0] if lo <= 0 <= hi else []])
    prog = single_op("Builtin1", 1, FNOP[fn], 0, n_in=1)
    ref = OracleProgram.from_program(prog).eval_batch([x], x.size)[0]
    got = gpu_eval(prog, [x], x.size)[0]
    d = ulp_distance(got, ref)
    assert d.max() <= 4, f"{fn}: {d.max()} ulp at x={x[np.argmax(d)]!r} (got {got[np.argmax(d)]!r}, ref {ref[np.argmax(d)]!r})"


def test_sign_of_zero_is_the_reference_s():
    z = np.array([0.0, -0.0])
    for op in ("Sin", "Exp", "Sqrt", "Neg", "Cos", "Square", "Cube", "Copy"):
        prog = single_op(op, 1, 0, n_in=1)
        ref = OracleProgram.from_program(prog).eval_batch([z], 2)[0]
        got = gpu_eval(prog, [z], 2)[0]
        assert np.array_equal(got, ref) and np.array_equal(np.signbit(got), np.signbit(ref)), op
    prog = single_op("SinCos", 1, 2, 0, n_in=1, n_out=2)
    ref = OracleProgram.from_program(prog).eval_batch([z], 2)
    got = gpu_eval(prog, [z], 2)
    for g, r in zip(got, ref):
        assert np.array_equal(g, r) and np.array_equal(np.signbit(g), np.signbit(r))
    for fn in ("tan", "asin", "atan", "sinh", "tanh", "expm1", "log1p", "cbrt", "asinh", "atanh", "erf", "floor", "ceil",
               "round", "abs", "signum"):
        prog = single_op("Builtin1", 1, FNOP[fn], 0, n_in=1)
        ref = OracleProgram.from_program(prog).eval_batch([z], 2)[0]
        got = gpu_eval(prog, [z], 2)[0]
        assert np.array_equal(got, ref) and np.array_equal(np.signbit(got), np.signbit(ref)), fn
    # -0 through arithmetic: (-0) + (-0), (-0) * 5, fma(-0, 1, -0)
    for op, cols in (("Add", [[-0.0], [-0.0]]), ("Mul", [[-0.0], [5.0]]), ("Sub", [[0.0], [0.0]])):
        prog = single_op(op, 2, 0, 1, n_in=2)
        c = [np.array(v) for v in cols]
        ref = OracleProgram.from_program(prog).eval_batch(c, 1)[0]
        got = gpu_eval(prog, c, 1)[0]
        assert np.array_equal(np.signbit(got), np.signbit(ref)) and got[0] == ref[0], op


# src/tests/fuzz_evaluator.rs:254-291: the SIMD instruction surface (one expression per builtin family)
SIMD_SURFACE = [
    ("asin(x/10)", ["x"]), ("acos(x/10)", ["x"]), ("atan(x)", ["x"]), ("asinh(x)", ["x"]), ("acosh(abs(x)+1.1)", ["x"]),
    ("atanh(x/10)", ["x"]), ("erf(x)", ["x"]), ("erfc(x)", ["x"]), ("gamma(abs(x)+1.2)", ["x"]), ("lgamma(abs(x)+1.2)", ["x"]),
    ("digamma(abs(x)+1.2)", ["x"]), ("trigamma(abs(x)+1.2)", ["x"]), ("tetragamma(abs(x)+1.2)", ["x"]),
    ("lambertw(x/10)", ["x"]), ("elliptic_k(x/2)", ["x"]), ("elliptic_e(x/2)", ["x"]), ("zeta(abs(x)+2.0)", ["x"]),
    ("exp_polar(x)", ["x"]), ("cbrt(x)", ["x"]), ("signum(x)", ["x"]), ("floor(x)", ["x"]), ("ceil(x)", ["x"]),
    ("round(x)", ["x"]), ("besselj(2, x)", ["x"]), ("bessely(2, abs(x)+0.5)", ["x"]), ("besseli(2, x)", ["x"]),
    ("besselk(2, abs(x)+0.5)", ["x"]), ("polygamma(2, abs(x)+1.2)", ["x"]), ("beta(abs(x)+1.1, abs(y)+1.2)", ["x", "y"]),
    ("zeta_deriv(1, abs(x)+2.0)", ["x"]), ("hermite(5, x)", ["x"]), ("log(2, abs(x)+1.1)", ["x"]), ("atan2(y, x)", ["x", "y"]),
    ("assoc_legendre(3, 1, x/2)", ["x"]), ("spherical_harmonic(3, 1, abs(x)+0.2, y)", ["x", "y"]),
    ("ynm(3, 1, abs(x)+0.2, y)", ["x", "y"]),
]


@pytest.mark.parametrize("expr,names", SIMD_SURFACE, ids=[e for e, _ in SIMD_SURFACE])
def test_reference_simd_surface_corpus(expr, names):
    rng = np.random.default_rng(sum(map(ord, expr)))
    n = 2048  # the reference draws 16 points per expression; more points, same range
    cols = [rng.uniform(-4.0, 4.0, n) for _ in names]
    prog = compile_exprs([parse(expr)], names)
    op = OracleProgram.from_program(prog)
    ref = op.eval_batch(cols, n)[0]
    got = gpu_eval(prog, cols, n)[0]
    # the reference's own bar between its scalar and SIMD paths is 1e-6 relative (fuzz_evaluator.rs:211-230); here the
    # end-to-end bar of the special functions (conditioning-amplified libm differences)
    rtol = 1e-9 if any(k in expr for k in ("bessel", "beta", "ynm", "spherical", "zeta_deriv")) else 2e-11
    assert_close(got, ref, rtol=rtol, atol=1e-12, what=expr)
    assert_close(op.run_chunked_simd(cols, n)[0], ref, rtol=0.0, atol=0.0, what=expr + " (oracle f64x4 body vs scalar)")
