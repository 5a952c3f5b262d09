"""Shared helpers of the test-suite: comparators, ulp distance, random program generator."""
from __future__ import annotations

import numpy as np

from symbanafis_b200.isa import FNOP, FNOP_ARITY, FNOP_NAME, Program, assemble


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-identical, except that any NaN matches any NaN (payloads differ between x86 and the GPU)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return bool(np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64)))


def ulp_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place; 0 for NaN/NaN and equal infinities, inf on class mismatch.
    Computed in exact integer arithmetic (a float64 cannot hold the difference of two 63-bit keys)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)

    def ordered(x):  # monotone map float64 -> int64 (two's complement ordering, -0.0 == +0.0)
        i = x.view(np.int64)
        return np.where(i < 0, np.int64(-(2 ** 63)) - i, i)

    ia, ib = ordered(a), ordered(b)
    same_sign = (ia < 0) == (ib < 0)
    with np.errstate(over="ignore"):
        d_int = np.abs(ia - ib)                      # exact when both keys have the same sign
    d_mixed = np.abs(ia.astype(np.float64)) + np.abs(ib.astype(np.float64))  # huge anyway
    d = np.where(same_sign, d_int.astype(np.float64), d_mixed)
    both_nan = np.isnan(a) & np.isnan(b)
    mismatch = np.isnan(a) ^ np.isnan(b)
    d = np.where(both_nan, 0.0, d)
    d = np.where(mismatch, np.inf, d)
    return d


def close_enough(got: np.ndarray, ref: np.ndarray, rtol: float = 1e-12, atol: float = 0.0) -> np.ndarray:
    """Elementwise: NaN==NaN, infinities must match in sign, else |got-ref| <= rtol*|ref| + atol
    (the reference's own comparator has the same shape, src/tests/fuzz_evaluator.rs:211-230)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    nan_ok = np.isnan(got) & np.isnan(ref)
    inf_ok = np.isinf(got) & np.isinf(ref) & (np.signbit(got) == np.signbit(ref))
    with np.errstate(invalid="ignore"):
        fin = np.isfinite(got) & np.isfinite(ref) & (np.abs(got - ref) <= rtol * np.abs(ref) + atol)
    return nan_ok | inf_ok | fin


def assert_close(got, ref, rtol=1e-12, atol=0.0, what=""):
    ok = close_enough(got, ref, rtol, atol)
    if not ok.all():
        i = int(np.flatnonzero(~ok)[0])
        raise AssertionError(f"{what}: {int((~ok).sum())} of {ok.size} differ; first at {i}: "
                             f"got {np.asarray(got)[i]!r} ref {np.asarray(ref)[i]!r}")


UNARY_OPS = ["Neg", "Square", "Cube", "Pow4", "Pow3_2", "InvPow3_2", "InvSqrt", "InvSquare", "InvCube", "Recip",
             "Sin", "Cos", "Exp", "Ln", "Sqrt", "RecipExpm1", "ExpSqr", "ExpSqrNeg", "Copy"]
BINARY_OPS = ["Add", "Mul", "Sub", "Div", "Pow", "NegMul"]
TERNARY_OPS = ["MulAdd", "MulSub", "NegMulAdd", "NegMulSub", "Add3", "Mul3"]
# builtins that are cheap enough and well-conditioned for random fuzzing
FUZZ_BUILTIN1 = ["tan", "cot", "sec", "csc", "asin", "acos", "atan", "acot", "asec", "acsc", "sinh", "cosh", "tanh",
                 "coth", "sech", "csch", "asinh", "acosh", "atanh", "acoth", "acsch", "asech", "expm1", "exp_neg",
                 "log1p", "cbrt", "abs", "signum", "floor", "ceil", "round", "erf", "erfc", "gamma", "lgamma",
                 "digamma", "trigamma", "tetragamma", "sinc", "lambert_w", "elliptic_k", "elliptic_e", "zeta",
                 "exp_polar"]
FUZZ_BUILTIN2 = ["atan2", "log", "bessel_j", "bessel_y", "bessel_i", "bessel_k", "polygamma", "beta", "hermite"]


def random_program(rng: np.random.Generator, n_params: int = 3, n_instr: int = 24, n_results: int = 2,
                   builtins: bool = True, n_consts: int = 4) -> Program:
    """A random but valid program in the reference ISA exercising every operand shape,
    including register reuse, in-place destinations, AddN/MulN pools and multiple results."""
    consts = rng.uniform(-2.0, 2.0, n_consts).round(3)
    base = n_params + n_consts
    n_temps = 10
    defined = list(range(base))  # params + consts readable from the start
    temps_defined: list = []
    instrs = []
    pool: list = []

    def src():
        return int(rng.choice(defined + temps_defined))

    def dst():
        return base + int(rng.integers(0, n_temps))

    for _ in range(n_instr):
        kind = rng.integers(0, 10 if builtins else 8)
        if kind <= 2:
            instrs.append((str(rng.choice(UNARY_OPS)), d := dst(), src()))
        elif kind <= 4:
            instrs.append((str(rng.choice(BINARY_OPS)), d := dst(), src(), src()))
        elif kind == 5:
            instrs.append((str(rng.choice(TERNARY_OPS)), d := dst(), src(), src(), src()))
        elif kind == 6:
            choice = rng.integers(0, 4)
            if choice == 0:
                instrs.append((str(rng.choice(["Add4", "Mul4"])), d := dst(), src(), src(), src(), src()))
            elif choice == 1:
                cnt = int(rng.integers(1, 7))
                start = len(pool)
                pool.extend(src() for _ in range(cnt))
                instrs.append((str(rng.choice(["AddN", "MulN"])), d := dst(), start, cnt))
            elif choice == 2:
                instrs.append(("Powi", d := dst(), src(), int(rng.integers(-6, 7))))
            else:
                d = dst()
                d2 = dst()
                instrs.append(("SinCos", d, d2, src()))
                if d2 not in temps_defined:
                    temps_defined.append(d2)
        elif kind == 7:
            instrs.append((str(rng.choice(["Sin", "Cos", "Exp"])), d := dst(), src()))
        elif kind == 8:
            instrs.append(("Builtin1", d := dst(), FNOP[str(rng.choice(FUZZ_BUILTIN1))], src()))
        else:
            instrs.append(("Builtin2", d := dst(), FNOP[str(rng.choice(FUZZ_BUILTIN2))], src(), src()))
        if d not in temps_defined:
            temps_defined.append(d)
    results = [int(rng.choice(temps_defined)) for _ in range(n_results)]
    return Program(flat=assemble(instrs), consts=consts, arg_pool=np.asarray(pool, dtype=np.uint32),
                   param_count=n_params, workspace_size=base + n_temps, result_regs=results,
                   param_names=[f"p{i}" for i in range(n_params)])
