#!/usr/bin/env python
"""Derives the constants of symbanafis_b200/csrc/leanmath.cuh (60-digit arithmetic, mpmath).

  sin:  P(z) ~ (sin r - r) / r^3, z = r^2, Chebyshev-node interpolation on |r| <= 1.0002 pi/2, 8 coefficients
  exp:  Q(r) ~ (exp r - 1 - r) / r^2 on |r| <= 1.0002 ln2/2, 11 coefficients
  pi split into three positive doubles (the middle part rounded down) for the Cody-Waite reduction
Run: python tools/fit_leanmath.py   (prints C initialisers; accuracy is then measured by leanmath_check.c)
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_fit(f, a, b, n):
    return list(reversed(mp.chebyfit(f, [a, b], n)))  # ascending coefficients


def main():
    Z = (mp.pi / 2 * mp.mpf("1.0002")) ** 2

    def fs(z):
        if z < mp.mpf("1e-30"):
            return -mp.mpf(1) / 6
        r = mp.sqrt(z)
        return (mp.sin(r) - r) / (z * r)

    c = cheb_fit(fs, 0, Z, 8)
    err = max(abs(mp.polyval(list(reversed(c)), z) - fs(z)) for z in mp.linspace(0, Z, 2000))
    print("sin: max |P - f| * z =", mp.nstr(err * Z, 5))
    print("kS = {" + ", ".join("%.17e" % float(x) for x in c) + "}")

    L = mp.log(2) / 2 * mp.mpf("1.0002")

    def fe(r):
        if abs(r) < mp.mpf("1e-25"):
            return mp.mpf(1) / 2
        return (mp.exp(r) - 1 - r) / (r * r)

    c = cheb_fit(fe, -L, L, 11)
    err = max(abs(mp.polyval(list(reversed(c)), r) - fe(r)) for r in mp.linspace(-L, L, 2001))
    print("exp: max |Q - f| * r^2 =", mp.nstr(err * L * L, 5))
    print("kE = {" + ", ".join("%.17e" % float(x) for x in c) + "}")

    # sincos with one reduction (leanmath.cuh sincos_fast): |r| <= 1.0002 pi/4, degree-5 S and C in z
    Zq = (mp.pi / 4 * mp.mpf("1.0002")) ** 2

    def fc(z):
        if z < mp.mpf("1e-20"):
            return mp.mpf(1) / 24
        r = mp.sqrt(z)
        return (mp.cos(r) - 1 + z / 2) / (z * z)

    c = cheb_fit(fs, 0, Zq, 6)
    err = max(abs(mp.polyval(list(reversed(c)), z) - fs(z)) for z in mp.linspace(0, Zq, 2000))
    print("sincos S: max |S - f| * z =", mp.nstr(err * Zq, 5))
    print("kQuad[4..9] = {" + ", ".join("%.17e" % float(x) for x in c) + "}")
    c = cheb_fit(fc, 0, Zq, 6)
    err = max(abs(mp.polyval(list(reversed(c)), z) - fc(z)) for z in mp.linspace(0, Zq, 2000))
    print("sincos C: max |C - f| * z^2 =", mp.nstr(err * Zq * Zq, 5))
    print("kQuad[10..15] = {" + ", ".join("%.17e" % float(x) for x in c) + "}")
    h = mp.pi / 2
    ha = float(h)
    hb = float(np.nextafter(float(h - mp.mpf(ha)), 0.0))
    hc = float(h - mp.mpf(ha) - mp.mpf(hb))
    print("kQuad[0..3] = {%.17e, %.17e, %.17e, %.17e}" % (float(2 / mp.pi), ha, hb, hc))

    pa = float(mp.pi)
    pb = float(np.nextafter(float(mp.pi - mp.mpf(pa)), 0.0))
    pc = float(mp.pi - mp.mpf(pa) - mp.mpf(pb))
    print("kPiA %.17e kPiB %.17e kPiC %.17e  kInvPi %.17e" % (pa, pb, pc, float(1 / mp.pi)))


if __name__ == "__main__":
    main()
