#!/usr/bin/env python
"""Generates tests/golden/reference_kats.json from the reference's OWN tests (run in the build container,
where /root/reference exists; the GPU box only sees the committed JSON).

The reference is a Rust crate and cannot be executed here (no cargo/rustc), so the oracle is pinned on the
known answers the reference's test-suite asserts at the boundary of this path:

* src/tests/precision_audit.rs        -- every `check_rel_error(eval("<expr>"), <value>, <TOL>, ..)` line:
                                         std functions to rel 1e-12, gamma family 1e-8, Bessel 1e-7
                                         (tolerance constants precision_audit.rs:6-9, comparison rule :32-57)
* src/evaluator/logic/bytecode/execute/drivers/tests.rs:10-90   -- eval_parallel! known answers
* tests/python/test_compiled_evaluator.py:27-193                -- CompiledEvaluator.evaluate / eval_batch
* tests/python/test_parallel_eval.py:19-70                      -- evaluate_parallel
* tests/python/test_numpy_returns.py:6-41                       -- eval_f64

The first group is extracted mechanically (regex over the Rust source, constants translated); the other
groups are short enough to be transcribed by hand below, each with its file:line.
"""
import json
import math
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")

TOL = {"REL_TOL": 1e-12, "REL_TOL_BESSEL": 1e-7, "REL_TOL_GAMMA": 1e-8}
ABS_TOL = 1e-14
EULER_GAMMA = 0.5772156649015329  # precision_audit.rs:350

RUST_CONSTS = {
    "std::f64::consts::E.powi(2)": math.e * math.e,
    "std::f64::consts::FRAC_1_SQRT_2": 0.7071067811865476,
    "std::f64::consts::FRAC_PI_2": math.pi / 2,
    "std::f64::consts::FRAC_PI_4": math.pi / 4,
    "std::f64::consts::LN_10": math.log(10.0),
    "std::f64::consts::LN_2": math.log(2.0),
    "std::f64::consts::SQRT_2": math.sqrt(2.0),
    "std::f64::consts::E": math.e,
    "PI.sqrt()": math.sqrt(math.pi),
    "EULER_GAMMA": EULER_GAMMA,
}


def rust_value(text: str):
    t = re.sub(r"\s+", " ", text.strip())
    for k in sorted(RUST_CONSTS, key=len, reverse=True):
        t = t.replace(k, repr(RUST_CONSTS[k]))
    t = re.sub(r"(?<=\d)_(?=\d)", "", t)
    if not re.fullmatch(r"[-+*/ .0-9e()]+", t):
        return None  # depends on a local variable: not a literal known answer
    return float(eval(t, {"__builtins__": {}}))  # arithmetic on literals only (checked by the regex above)


def line_of(src: str, pos: int) -> int:
    return src.count("\n", 0, pos) + 1


def precision_audit():
    path = os.path.join(REF, "src/tests/precision_audit.rs")
    src = open(path).read()
    pat = re.compile(r'check_rel_error\(\s*eval\("([^"]+)"\),\s*(.+?),\s*(REL_TOL\w*),\s*"', re.S)
    cases = []
    for m in pat.finditer(src):
        v = rust_value(m.group(2))
        if v is None:
            continue
        cases.append({"expr": m.group(1), "params": [], "inputs": [], "expected": v,
                      "rel_tol": TOL[m.group(3)], "abs_tol_near_zero": ABS_TOL,
                      "source": f"src/tests/precision_audit.rs:{line_of(src, m.start())}"})
    return cases


def transcribed():
    e10 = 1e-10
    c = []

    def scalar(expr, params, inputs, expected, source, tol=e10):
        c.append({"kind": "evaluate", "expr": expr, "params": params, "inputs": inputs, "expected": expected,
                  "tol": tol, "source": source})

    def batch(expr, params, columns, expected, source, tol=e10):
        c.append({"kind": "eval_batch", "expr": expr, "params": params, "columns": columns, "expected": expected,
                  "tol": tol, "source": source})

    py = "tests/python/test_compiled_evaluator.py"
    scalar("x^2 + 1", ["x"], [2.0], 5.0, f"{py}:27-32")
    scalar("x + y", ["x", "y"], [3.0, 4.0], 7.0, f"{py}:34-39")
    scalar("sin(x)", ["x"], [math.pi / 2], 1.0, f"{py}:41-46")
    scalar("exp(x)", ["x"], [1.0], math.e, f"{py}:48-53")
    batch("x^2", ["x"], [[1.0, 2.0, 3.0, 4.0]], [1.0, 4.0, 9.0, 16.0], f"{py}:59-70")
    batch("x * y", ["x", "y"], [[1.0, 2.0, 3.0], [10.0, 20.0, 30.0]], [10.0, 40.0, 90.0], f"{py}:72-82")
    batch("x^2 + x", ["x"], [[1.0, 2.0, 3.0, 4.0, 5.0]], [2.0, 6.0, 12.0, 20.0, 30.0], f"{py}:84-93")
    scalar("besselj(0, x)", ["x"], [0.0], 1.0, f"{py}:112-117")
    scalar("gamma(x)", ["x"], [5.0], 24.0, f"{py}:119-124")
    scalar("erf(x)", ["x"], [0.0], 0.0, f"{py}:126-131")
    for x in [0.0, 0.5, 1.0, 2.0, 3.14159]:
        scalar("sin(x)^2 + cos(x)^2", ["x"], [x], 1.0, f"{py}:133-140")
    scalar("x^2 + x*3 + 1", ["x"], [2.0], 11.0, f"{py}:146-152")
    scalar("x*y + x + y", ["x", "y"], [2.0, 3.0], 11.0, f"{py}:154-161")
    scalar("sin(x) + cos(x)", ["x"], [0.0], 1.0, f"{py}:163-169")
    scalar("42", [], [], 42.0, f"{py}:175-180")
    scalar("sin(pi/2)", [], [], 1.0, f"{py}:182-187")
    scalar("ln(e)", [], [], 1.0, f"{py}:189-193")

    dr = "src/evaluator/logic/bytecode/execute/drivers/tests.rs"
    batch("x^2", ["x"], [[1.0, 2.0, 3.0]], [1.0, 4.0, 9.0], f"{dr}:10-25", tol=0.0)
    batch("x^2", ["x"], [[2.0, 3.0]], [4.0, 9.0], f"{dr}:48-66", tol=0.0)
    batch("t + 1", ["t"], [[10.0, 20.0]], [11.0, 21.0], f"{dr}:48-72")
    batch("x + y", ["x", "y"], [[1.0, 2.0], [10.0, 20.0]], [11.0, 22.0], f"{dr}:74-88", tol=0.0)
    batch("x * y", ["x", "y"], [[2.0], [3.0]], [6.0], f"{dr}:90-104 (numeric point of the SKIP test)", tol=0.0)

    pe = "tests/python/test_parallel_eval.py"
    batch("x^2", ["x"], [[1.0, 2.0, 3.0]], [1.0, 4.0, 9.0], f"{pe}:19-30", tol=0.0)
    batch("x + y", ["x", "y"], [[1.0, 2.0], [10.0, 20.0]], [11.0, 22.0], f"{pe}:32-42", tol=0.0)
    batch("x + 1", ["x"], [[1.0, 2.0, 3.0]], [2.0, 3.0, 4.0], f"{pe}:44-56", tol=0.0)
    batch("x^2", ["x"], [[1.0, 2.0, 3.0, 4.0]], [1.0, 4.0, 9.0, 16.0], f"{pe}:62-71", tol=0.0)

    nr = "tests/python/test_numpy_returns.py"
    batch("x^2", ["x"], [[1.0, 2.0, 3.0]], [1.0, 4.0, 9.0], f"{nr}:6-41")
    batch("x + y", ["x", "y"], [[1.0, 2.0, 3.0], [10.0, 20.0, 30.0]], [11.0, 22.0, 33.0], f"{nr}:6-41")
    return c


def main():
    doc = {
        "generated_by": "tests/golden/make_golden.py (reads the reference's tests; never executed on the GPU box)",
        "comparison_rule": "precision_audit: |c-e| < abs_tol_near_zero when |e| < abs_tol_near_zero, else |c-e|/|e| < rel_tol "
                           "(precision_audit.rs:32-57); transcribed: |a-b| < tol*max(|a|,|b|,1) "
                           "(test_compiled_evaluator.py:16-22), tol 0 = exact",
        "precision_audit": precision_audit(),
        "transcribed": transcribed(),
    }
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1)
    print(f"wrote {OUT}: {len(doc['precision_audit'])} precision_audit cases, {len(doc['transcribed'])} transcribed")


if __name__ == "__main__":
    main()
