"""Derives the double-pendulum angular accelerations a1, a2 from the Lagrangian (sympy) and writes
them as expression strings to symbanafis_b200/data/double_pendulum_rhs.json.

Same derivation as the reference example (examples/python/double_pendulum_benchmark.py:39-108,
derive_sympy): Lagrangian of two point masses, mass matrix from d2L/dw2, Cramer's rule, then
m1 = m2 = L1 = L2 = 1, g = 9.81 (:32-35).  The strings are what `parse()` consumes.
"""
import json
import os

import sympy as sp

G, L1, L2, M1, M2 = 9.81, 1.0, 1.0, 1.0, 1.0


def derive():
    th1, th2, w1, w2 = sp.symbols("th1 th2 w1 w2")
    m1, m2, l1, l2, g = sp.symbols("m1 m2 L1 L2 g")
    x1 = l1 * sp.sin(th1)
    y1 = -l1 * sp.cos(th1)
    y2 = y1 - l2 * sp.cos(th2)
    vx1 = l1 * w1 * sp.cos(th1)
    vy1 = l1 * w1 * sp.sin(th1)
    vx2 = vx1 + l2 * w2 * sp.cos(th2)
    vy2 = vy1 + l2 * w2 * sp.sin(th2)
    T = sp.Rational(1, 2) * m1 * (vx1**2 + vy1**2) + sp.Rational(1, 2) * m2 * (vx2**2 + vy2**2)
    V = m1 * g * y1 + m2 * g * y2
    L = T - V
    dL_dw1, dL_dw2 = sp.diff(L, w1), sp.diff(L, w2)
    dL_dth1, dL_dth2 = sp.diff(L, th1), sp.diff(L, th2)
    M11, M12 = sp.diff(dL_dw1, w1), sp.diff(dL_dw1, w2)
    M21, M22 = sp.diff(dL_dw2, w1), sp.diff(dL_dw2, w2)
    dt1 = sp.diff(dL_dw1, th1) * w1 + sp.diff(dL_dw1, th2) * w2
    dt2 = sp.diff(dL_dw2, th1) * w1 + sp.diff(dL_dw2, th2) * w2
    F1, F2 = dL_dth1 - dt1, dL_dth2 - dt2
    det = M11 * M22 - M12 * M21
    a1 = (F1 * M22 - F2 * M12) / det
    a2 = (M11 * F2 - M21 * F1) / det
    subs = {m1: M1, m2: M2, l1: L1, l2: L2, g: G}
    a1 = sp.simplify(a1.subs(subs))
    a2 = sp.simplify(a2.subs(subs))
    return str(a1), str(a2)


if __name__ == "__main__":
    a1, a2 = derive()
    out = os.path.join(os.path.dirname(__file__), "..", "symbanafis_b200", "data", "double_pendulum_rhs.json")
    with open(out, "w") as f:
        json.dump({"params": ["th1", "th2", "w1", "w2"], "a1": a1, "a2": a2,
                   "source": "tools/derive_double_pendulum.py (sympy %s)" % sp.__version__}, f, indent=1)
    print(a1)
    print(a2)
