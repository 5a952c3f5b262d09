/* leanmath_check.c -- CPU accuracy check of the kernels in symbanafis_b200/csrc/leanmath.cuh.
 *
 * The device code is a sequence of correctly rounded fma/mul/add operations and integer bit
 * manipulation, so the same sequence with C99 fma() on the CPU produces the SAME bits; this file
 * restates the sequences (keep in sync with leanmath.cuh) and measures their error in ulp against
 * the x87 80-bit long double libm (error of the reference <= 2^-11 double ulp) and against glibc's
 * double functions (what the reference evaluator calls on Linux).
 *
 *   gcc -O2 -ffp-contract=off -mfma -o /tmp/leanmath_check tools/leanmath_check.c -lm && /tmp/leanmath_check
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const double MAGIC = 6755399441055744.0; /* 1.5 * 2^52 */
static const double INV_PI = 3.18309886183790691e-01;
static const double PI_A = 3.14159265358979312e+00, PI_B = 1.22464679914735296e-16, PI_C = 2.16571334784382806e-32; /* all parts > 0: sin(-0) = -0 */
static const double S8[] = {-1.66666666666666657e-01, 8.33333333333331587e-03, -1.98412698412549389e-04,
                            2.75573192191536246e-06,  -2.50521076157725835e-08, 1.60589772331473175e-10,
                            -7.64396776493426846e-13, 2.73141320376900007e-15};
static const double LOG2E = 1.44269504088896338700e+00, LN2_HI = 6.93147180369123816490e-01,
                    LN2_LO = 1.90821492927058770002e-10;
static const double E11[] = {5.00000000000000000e-01, 1.66666666666666713e-01, 4.16666666666666713e-02,
                             8.33333333332612891e-03, 1.38888888888837438e-03, 1.98412698748407683e-04,
                             2.48015873255621253e-05, 2.75572553746587440e-06, 2.75572736248664783e-07,
                             2.51052276365517599e-08, 2.09146945603515922e-09};

static uint64_t bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static double from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

static double sin_kernel(double r, uint32_t k) { /* (-1)^k sin(r), |r| <= pi/2 */
    r = from_bits(bits(r) ^ ((uint64_t)(k & 1u) << 63)); /* sin is odd: flip the argument */
    const double z = r * r;
    double p = S8[7];
    for (int i = 6; i >= 0; --i) p = fma(p, z, S8[i]);
    const double v = fma(r * z, p, r);
    /* the result has the sign of r (also for r = -0, where the fma above yields +0) */
    return from_bits((bits(v) & 0x7fffffffffffffffull) | (bits(r) & 0x8000000000000000ull));
}
static double lean_sin(double x) {
    const double t = fma(x, INV_PI, MAGIC);
    const uint32_t k = (uint32_t)bits(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -PI_A, x);
    r = fma(kd, -PI_B, r);
    r = fma(kd, -PI_C, r);
    return sin_kernel(r, k);
}
static double lean_cos(double x) { /* cos(x) = (-1)^(k+1) sin(x - (k + 1/2) pi), k = rint(x/pi - 1/2) */
    const double t = fma(x, INV_PI, -0.5) + MAGIC;
    const uint32_t k = (uint32_t)bits(t);
    const double hd = (t - MAGIC) + 0.5;
    double r = fma(hd, -PI_A, x);
    r = fma(hd, -PI_B, r);
    r = fma(hd, -PI_C, r);
    return sin_kernel(r, k + 1u);
}
static double lean_exp(double x) {
    const double t = fma(x, LOG2E, MAGIC);
    const int32_t k = (int32_t)(uint32_t)bits(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -LN2_HI, x);
    r = fma(kd, -LN2_LO, r);
    double p = E11[10];
    for (int i = 9; i >= 0; --i) p = fma(p, r, E11[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return from_bits(bits(p) + ((uint64_t)(int64_t)k << 52));
}

/* sincos with ONE reduction (leanmath.cuh sincos_fast): k = rint(x * 2/pi), r = x - k pi/2, |r| <= pi/4; sin r and cos r
 * from two degree-5 polynomials in z = r^2; the quadrant swaps / negates them.  20 FP64 operations for both values. */
static const double TWO_OVER_PI = 6.36619772367581382e-01;
static const double PIO2_A = 1.57079632679489656e+00, PIO2_B = 6.12323399573676480e-17, PIO2_C = 1.08285667392191403e-32;
static const double QS[] = {-1.66666666666666657e-01, 8.33333333333094450e-03, -1.98412698367513636e-04,
                            2.75573160988126868e-06, -2.50511310657360491e-08, 1.59180731629230736e-10};
static const double QC[] = {4.16666666666666644e-02, -1.38888888888873932e-03, 2.48015872987611761e-05,
                            -2.75573172693902722e-07, 2.08761457809035442e-09, -1.13825973098618771e-11};
static void lean_sincos(double x, double *sn, double *cs) {
    const double t = fma(x, TWO_OVER_PI, MAGIC);
    const uint32_t q = (uint32_t)bits(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -PIO2_A, x);
    r = fma(kd, -PIO2_B, r);
    r = fma(kd, -PIO2_C, r);
    const double z = r * r;
    double ps = QS[5], pc = QC[5];
    for (int i = 4; i >= 0; --i) { ps = fma(ps, z, QS[i]); pc = fma(pc, z, QC[i]); }
    double s = fma(r * z, ps, r);
    s = from_bits((bits(s) & 0x7fffffffffffffffull) | (bits(r) & 0x8000000000000000ull)); /* sign of r, also for -0 */
    const double c = fma(z, fma(z, pc, -0.5), 1.0);
    const double a = (q & 1u) ? c : s, b = (q & 1u) ? s : c; /* odd quadrant: swap */
    *sn = from_bits(bits(a) ^ ((uint64_t)((q >> 1) & 1u) << 63));
    *cs = from_bits(bits(b) ^ ((uint64_t)(((q + 1u) >> 1) & 1u) << 63));
}

static double ulp_of(double ref) {
    int e;
    frexp(ref, &e);
    return ldexp(1.0, e - 53);
}
static double ulp_err(double got, long double ref) {
    const double rd = (double)ref;
    if (rd == 0.0) return got == 0.0 ? 0.0 : INFINITY;
    return (double)fabsl(((long double)got - ref) / (long double)ulp_of(rd));
}
static double rnd(uint64_t *s) { /* xorshift, uniform in [0,1) */
    *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17;
    return (double)(*s >> 11) * (1.0 / 9007199254740992.0);
}

int main(void) {
    uint64_t s = 0x9e3779b97f4a7c15ull;
    const double ranges[][2] = {{0, 1e-3}, {0, 1.6}, {0, 7}, {0, 100}, {0, 1e4}, {0, 1048576.0}};
    int fail = 0;
    for (unsigned ri = 0; ri < sizeof ranges / sizeof ranges[0]; ++ri) {
        double ws = 0, wc = 0, gs = 0, gc = 0;
        for (int i = 0; i < 4000000; ++i) {
            double x = ranges[ri][0] + (ranges[ri][1] - ranges[ri][0]) * rnd(&s);
            if (i & 1) x = -x;
            const double e1 = ulp_err(lean_sin(x), sinl((long double)x));
            const double e2 = ulp_err(lean_cos(x), cosl((long double)x));
            const double g1 = fabs(lean_sin(x) - sin(x)) / ulp_of(sin(x));
            const double g2 = fabs(lean_cos(x) - cos(x)) / ulp_of(cos(x));
            if (e1 > ws) ws = e1;
            if (e2 > wc) wc = e2;
            if (g1 > gs) gs = g1;
            if (g2 > gc) gc = g2;
        }
        printf("|x| < %-10g  sin: %.3f ulp (true)  %.0f ulp vs glibc   cos: %.3f ulp (true)  %.0f ulp vs glibc\n",
               ranges[ri][1], ws, gs, wc, gc);
        if (ws > 2.5 || wc > 2.5) fail = 1;
    }
    for (unsigned ri = 0; ri < sizeof ranges / sizeof ranges[0]; ++ri) {
        double ws = 0, wc = 0, gs = 0, gc = 0;
        for (int i = 0; i < 4000000; ++i) {
            double x = ranges[ri][0] + (ranges[ri][1] - ranges[ri][0]) * rnd(&s);
            if (i & 1) x = -x;
            double sn, cs;
            lean_sincos(x, &sn, &cs);
            const double e1 = ulp_err(sn, sinl((long double)x)), e2 = ulp_err(cs, cosl((long double)x));
            const double g1 = fabs(sn - sin(x)) / ulp_of(sin(x)), g2 = fabs(cs - cos(x)) / ulp_of(cos(x));
            if (e1 > ws) ws = e1;
            if (e2 > wc) wc = e2;
            if (g1 > gs) gs = g1;
            if (g2 > gc) gc = g2;
        }
        printf("sincos |x| < %-10g  sin: %.3f ulp (true)  %.0f ulp vs glibc   cos: %.3f ulp (true)  %.0f ulp vs glibc\n",
               ranges[ri][1], ws, gs, wc, gc);
        if (ws > 2.5 || wc > 2.5) fail = 1;
    }
    {
        double ws2 = 0, wc2 = 0;
        for (int k = -700000; k <= 700000; ++k) {
            const double c = (double)((long double)k * 1.57079632679489661923132169163975144L);
            for (int d = -2; d <= 2; ++d) {
                double x = c, sn, cs;
                for (int j = 0; j < abs(d); ++j) x = nextafter(x, d > 0 ? INFINITY : -INFINITY);
                lean_sincos(x, &sn, &cs);
                const double e1 = ulp_err(sn, sinl((long double)x)), e2 = ulp_err(cs, cosl((long double)x));
                if (e1 > ws2) ws2 = e1;
                if (e2 > wc2) wc2 = e2;
            }
        }
        printf("sincos next to k*pi/2, |k| <= 7e5: sin %.3f ulp, cos %.3f ulp\n", ws2, wc2);
        if (ws2 > 2.5 || wc2 > 2.5) fail = 1;
        double sn, cs;
        lean_sincos(-0.0, &sn, &cs);
        if (!signbit(sn) || cs != 1.0) fail = 1;
        lean_sincos(5e-324, &sn, &cs);
        if (sn != 5e-324 || cs != 1.0) fail = 1;
    }
    /* points next to multiples of pi/2 (worst cancellation) */
    double ws = 0, wc = 0;
    for (int k = -200000; k <= 200000; ++k) {
        const double c = (double)((long double)k * 1.57079632679489661923132169163975144L);
        for (int d = -2; d <= 2; ++d) {
            double x = c;
            for (int j = 0; j < abs(d); ++j) x = nextafter(x, d > 0 ? INFINITY : -INFINITY);
            const double e1 = ulp_err(lean_sin(x), sinl((long double)x));
            const double e2 = ulp_err(lean_cos(x), cosl((long double)x));
            if (e1 > ws) ws = e1;
            if (e2 > wc) wc = e2;
        }
    }
    printf("next to k*pi/2, |k| <= 2e5: sin %.3f ulp, cos %.3f ulp\n", ws, wc);
    if (ws > 2.0 || wc > 2.0) fail = 1;
    /* special values of the fast path */
    printf("sin(+0)=%g sin(-0)=%g (signbit %d) sin(5e-324)=%g cos(0)=%.17g cos(-0)=%.17g\n", lean_sin(0.0), lean_sin(-0.0),
           signbit(lean_sin(-0.0)), lean_sin(5e-324), lean_cos(0.0), lean_cos(-0.0));
    if (!signbit(lean_sin(-0.0)) || lean_sin(5e-324) != 5e-324 || lean_cos(0.0) != 1.0) fail = 1;

    double we = 0, ge = 0;
    for (int i = 0; i < 8000000; ++i) {
        const double x = -700.0 + 1400.0 * rnd(&s);
        const double e = ulp_err(lean_exp(x), expl((long double)x));
        const double g = fabs(lean_exp(x) - exp(x)) / ulp_of(exp(x));
        if (e > we) we = e;
        if (g > ge) ge = g;
    }
    for (int i = 0; i < 4000000; ++i) {
        const double x = -1.0 + 2.0 * rnd(&s);
        const double e = ulp_err(lean_exp(x), expl((long double)x));
        if (e > we) we = e;
    }
    printf("exp on [-700, 700]: %.3f ulp (true), %.0f ulp vs glibc; exp(0)=%.17g exp(-0)=%.17g\n", we, ge, lean_exp(0.0),
           lean_exp(-0.0));
    if (we > 1.5 || lean_exp(0.0) != 1.0) fail = 1;
    printf(fail ? "FAIL\n" : "OK\n");
    return fail;
}
