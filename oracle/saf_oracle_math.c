/*
 * saf_oracle_math.c -- special functions of the SymbAnaFis evaluator, restated
 * in C for T = f64.  TEST INFRASTRUCTURE ONLY (see saf_oracle.h).
 *
 * Every function follows the named Rust source expression by expression, in
 * the same evaluation order; build with -ffp-contract=off so that no a*b+c is
 * fused (rustc never contracts).
 */
#include "saf_oracle_math.h"

#include <float.h>
#include <math.h>
#include <stdint.h>

#define SAF_PI 3.14159265358979323846264338327950288  /* f64::consts::PI */
#define SAF_E 2.71828182845904523536028747135266250   /* f64::consts::E */
#define SAF_FRAC_2_PI 0.636619772367581343075535053490057448 /* f64::consts::FRAC_2_PI */
#define SAF_EPSILON 1e-14                              /* crate::EPSILON, src/lib.rs:321 */

/* Work caps (documented deviation, DESIGN.md "Unbounded loops of the reference"): the reference walks
 * `while xv < 6 { xv += 1 }` recurrences and order-n loops without any bound, so trigamma(-inf) never
 * returns and besselj(2^31-1, x) runs for minutes.  Oracle and device both answer NaN instead when the
 * reference would need more than SAF_MAX_SHIFT recurrence steps or an order above SAF_MAX_ORDER. */
#define SAF_MAX_SHIFT 4194304.0
#define SAF_MAX_ORDER 65536

/* ---- Rust std / compiler-rt pieces that are not under /root/reference ---- */

/* f64::powi -> llvm.powi.f64.i32 -> compiler-rt __powidf2 (published algorithm:
 * square-and-multiply from the least significant bit, reciprocal at the end). */
double saf_o_powi(double a, int32_t b) {
    const int recip = b < 0;
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

/* f64::signum: 1.0 for +0.0 and positives, -1.0 for -0.0 and negatives, NaN for NaN. */
double saf_o_signum(double x) {
    if (isnan(x)) return x;
    return copysign(1.0, x);
}

/* Rust std f64::asinh / acosh / atanh (library/std/src/f64.rs). Parity unpinned:
 * the formulas are restated from the published std source, not from /root/reference. */
double saf_o_asinh(double x) {
    double ax = fabs(x);
    double ix = 1.0 / ax;
    return copysign(log1p(ax + (ax / (hypot(1.0, ix) + ix))), x);
}
double saf_o_acosh(double x) {
    if (x < 1.0) return NAN;
    return log(x + (sqrt(x - 1.0) * sqrt(x + 1.0)));
}
double saf_o_atanh(double x) { return 0.5 * log1p((2.0 * x) / (1.0 - x)); }

/* ---- helpers.rs:3-36 ---- */
static double sign_from_delta(double delta) {
    if (isnan(delta)) return NAN;
    return signbit(delta) ? -1.0 : 1.0;
}
static double signed_infinity(double sign) {
    if (isnan(sign)) return NAN;
    return signbit(sign) ? -INFINITY : INFINITY;
}
static double signed_infinity_from_delta(double delta) {
    return signed_infinity(sign_from_delta(delta));
}
/* num_traits ToPrimitive::to_i64 for f64, then unwrap_or(0) */
static int64_t to_i64_or0(double v) {
    if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return 0;
    return (int64_t)v;
}
static double gamma_pole_sign(double x) {
    int64_t n = to_i64_or0(-x);
    return (n & 1) == 0 ? 1.0 : -1.0;
}
static double fract(double x) { return x - trunc(x); }

/* ---- erf.rs:9-51 ---- */
double saf_o_erf(double x0) {
    double sign = saf_o_signum(x0);
    double x = fabs(x0);
    double sqrt_pi = sqrt(SAF_PI);
    double coeff = 2.0 / sqrt_pi;
    double sum = 0.0, compensation = 0.0, factorial = 1.0, power = x;
    for (int n = 0; n < 30; ++n) {
        double two_n_plus_one = (double)(2 * n + 1);
        double term = power / (factorial * two_n_plus_one);
        if (isnan(term) || isinf(term)) break;
        double signed_term = (n % 2 == 0) ? term : -term;
        double y = signed_term - compensation;
        double t = sum + y;
        compensation = (t - sum) - y;
        sum = t;
        factorial *= (double)(n + 1);
        power *= x * x;
        if (fabs(term) < DBL_EPSILON) break;
    }
    return sign * coeff * sum;
}

/* ---- erf.rs:64-126 ---- */
double saf_o_erfc(double x) {
    if (isnan(x)) return x;
    if (isinf(x)) return signbit(x) ? 2.0 : 0.0;
    double abs_x = fabs(x);
    if (x > 28.0) return 0.0;
    if (abs_x < 1.5) return 1.0 - saf_o_erf(x);
    if (x < 0.0) return 2.0 - saf_o_erfc(abs_x);

    const double tiny = 1e-30, eps = DBL_EPSILON;
    double f = abs_x;
    if (fabs(f) < tiny) f = tiny;
    double c = f, d = 0.0;
    for (int j = 1; j < 200; ++j) {
        double a = 0.5 * (double)j;
        d = abs_x + a * d;
        if (fabs(d) < tiny) d = tiny;
        d = 1.0 / d;
        c = abs_x + a / c;
        if (fabs(c) < tiny) c = tiny;
        double delta = c * d;
        f *= delta;
        if (fabs(delta - 1.0) < eps) break;
    }
    double coeff = exp(-abs_x * abs_x) / sqrt(SAF_PI);
    return coeff / f;
}

/* ---- gamma.rs:4-24 ---- */
static double lanczos_ag(double x_minus_one) {
    static const double c[9] = {
        0.9999999999998099, 676.5203681218851,     -1259.1392167224028,
        771.3234287776531,  -176.6150291621406,    12.507343278686905,
        -0.13857109526572012, 9.984369578019572e-6, 1.5056327351493116e-7,
    };
    double ag = c[0];
    for (int i = 1; i < 9; ++i) ag += c[i] / (x_minus_one + (double)i);
    return ag;
}

/* ---- gamma.rs:33-65 ---- */
double saf_o_gamma(double x) {
    if (x <= 0.0 && fract(x) == 0.0) {
        double delta = x - round(x);
        double sign = gamma_pole_sign(x) * sign_from_delta(delta);
        return signed_infinity(sign);
    }
    const double half = 0.5, one = 1.0, pi = SAF_PI;
    if (x < half) {
        return pi / (sin(pi * x) * saf_o_gamma(one - x));
    }
    double x_minus_one = x - one;
    double ag = lanczos_ag(x_minus_one);
    double g = 7.0;
    double t = x_minus_one + g + half;
    double two_pi_sqrt = sqrt(2.0 * pi);
    return two_pi_sqrt * pow(t, x_minus_one + half) * exp(-t) * ag;
}

/* ---- gamma.rs:70-77 ---- */
static double gamma_sign(double x) {
    if (x > 0.0) return 1.0;
    int64_t n = to_i64_or0(-x);
    return (n & 1) == 0 ? -1.0 : 1.0;
}

/* ---- gamma.rs:84-103 ---- */
double saf_o_lgamma(double x) {
    if (x <= 0.0 && fract(x) == 0.0) return INFINITY;
    const double half = 0.5, one = 1.0, pi = SAF_PI;
    if (x < half) {
        double sin_term = fabs(sin(pi * x));
        return log(pi) - log(sin_term) - saf_o_lgamma(one - x);
    }
    double x_minus_one = x - one;
    double ag = lanczos_ag(x_minus_one);
    double g = 7.0;
    double t = x_minus_one + g + half;
    double two_pi_sqrt_ln = log(sqrt(2.0 * pi));
    return two_pi_sqrt_ln + (x_minus_one + half) * log(t) - t + log(fabs(ag));
}

/* ---- beta.rs:11-33 ---- */
double saf_o_beta(double x, double y) {
    if (x <= 0.0 && fract(x) == 0.0) return signed_infinity(gamma_sign(x));
    if (y <= 0.0 && fract(y) == 0.0) return signed_infinity(gamma_sign(y));
    double xy = x + y;
    if (xy <= 0.0 && fract(xy) == 0.0) return 0.0;
    double sign = gamma_sign(x) * gamma_sign(y) * gamma_sign(xy);
    double lbeta = saf_o_lgamma(x) + saf_o_lgamma(y) - saf_o_lgamma(xy);
    return sign * exp(lbeta);
}

/* ---- polygamma.rs:11-41 ---- */
double saf_o_digamma(double x) {
    if (x <= 0.0 && fract(x) == 0.0) {
        double delta = x - round(x);
        double sign = -sign_from_delta(delta);
        return signed_infinity(sign);
    }
    const double half = 0.5, one = 1.0, pi = SAF_PI;
    if (x < half) {
        return saf_o_digamma(one - x) - pi * cos(pi * x) / sin(pi * x);
    }
    double xv = x, result = 0.0;
    const double six = 6.0;
    while (xv < six) {
        result -= one / xv;
        xv += one;
    }
    result += log(xv) - half / xv;
    double x2 = xv * xv;
    double t1 = one / (12.0 * x2);
    double t2 = one / (120.0 * x2 * x2);
    double t3 = one / (252.0 * x2 * x2 * x2);
    return result - t1 + t2 - t3;
}

/* ---- polygamma.rs:49-68 ---- */
double saf_o_trigamma(double x) {
    if (x <= 0.0 && fract(x) == 0.0) return INFINITY;
    if (x < -SAF_MAX_SHIFT) return NAN; /* work cap */
    double xv = x, r = 0.0;
    const double six = 6.0, one = 1.0;
    while (xv < six) {
        r += one / (xv * xv);
        xv += one;
    }
    double x2 = xv * xv;
    const double half = 0.5;
    return r + one / xv + half / x2 + one / (six * x2 * xv) - one / (30.0 * x2 * x2 * xv) +
           one / (42.0 * x2 * x2 * x2 * xv);
}

/* ---- polygamma.rs:75-94 ---- */
double saf_o_tetragamma(double x) {
    if (x <= 0.0 && fract(x) == 0.0) {
        double delta = x - round(x);
        double sign = -sign_from_delta(delta);
        return signed_infinity(sign);
    }
    if (x < -SAF_MAX_SHIFT) return NAN; /* work cap */
    double xv = x, r = 0.0;
    const double six = 6.0, one = 1.0, two = 2.0;
    while (xv < six) {
        r -= two / (xv * xv * xv);
        xv += one;
    }
    double x2 = xv * xv;
    return r - one / x2 + one / (x2 * xv) + one / (two * x2 * x2) + one / (six * x2 * x2 * xv);
}

/* ---- polygamma.rs:102-208 ---- */
double saf_o_polygamma(int32_t n, double x) {
    if (n < 0) return NAN;
    if (n == 0) return saf_o_digamma(x);
    if (n == 1) return saf_o_trigamma(x);
    if (n > SAF_MAX_ORDER || x < -SAF_MAX_SHIFT) return NAN; /* work cap */

    if (x <= 0.0 && fract(x) == 0.0) {
        double delta = x - round(x);
        double sign = (n % 2 == 0) ? -sign_from_delta(delta) : 1.0;
        return signed_infinity(sign);
    }
    double xv = x, r = 0.0;
    double sign = ((n + 1) % 2 == 0) ? 1.0 : -1.0;
    double factorial = 1.0;
    for (int32_t i = 1; i <= n; ++i) factorial *= (double)i;
    const double fifteen = 15.0, one = 1.0;
    int32_t n_plus_one = n + 1;
    while (xv < fifteen) {
        r += sign * factorial / saf_o_powi(xv, n_plus_one);
        xv += one;
    }
    double asym_sign = (n % 2 == 0) ? -1.0 : 1.0;
    static const double bernoulli_pairs[5][2] = {
        {1.0, 6.0}, {-1.0, 30.0}, {1.0, 42.0}, {-1.0, 30.0}, {5.0, 66.0}};
    double n_minus_1_fact = 1.0;
    if (n > 1)
        for (int32_t i = 1; i < n; ++i) n_minus_1_fact *= (double)i;

    double sum = n_minus_1_fact / saf_o_powi(xv, n);
    const double two = 2.0;
    sum += factorial / (two * saf_o_powi(xv, n_plus_one));

    double xpow = saf_o_powi(xv, n + 2);
    double fact_ratio = factorial * (double)(n + 1);
    double prev_term_abs = DBL_MAX;
    for (int k = 0; k < 5; ++k) {
        int32_t k_plus_1 = k + 1;
        int32_t two_k = 2 * k_plus_1;
        double factorial_2k = 1.0;
        for (int32_t i = 1; i <= two_k; ++i) factorial_2k *= (double)i;
        double val_bk = bernoulli_pairs[k][0] / bernoulli_pairs[k][1];
        double term = val_bk * fact_ratio / (factorial_2k * xpow);
        if (fabs(term) > prev_term_abs) break;
        prev_term_abs = fabs(term);
        sum += term;
        xpow *= xv * xv;
        double next_factor1 = (double)(n + two_k);
        double next_factor2 = (double)(n + two_k + 1);
        fact_ratio *= next_factor1 * next_factor2;
    }
    return r + asym_sign * sum;
}

/* ---- zeta.rs:107-163 ---- */
static double zeta_borwein(double s) {
    const double one = 1.0, two = 2.0, four = 4.0;
    enum { N = 14 };
    double denom = one - pow(two, one - s);
    if (fabs(denom) < 1e-15) {
        double delta = s - one;
        return signed_infinity_from_delta(delta);
    }
    double d_coeffs[N + 1];
    double n_t = (double)N;
    double term = one / n_t;
    double current_inner_sum = term;
    d_coeffs[0] = n_t * current_inner_sum;
    for (int k = 1; k <= N; ++k) {
        double i = (double)(k - 1);
        double two_i_plus_1 = (double)(2 * k - 1);
        double two_i_plus_2 = (double)(2 * k);
        double n_minus_i = n_t - i;
        double n_plus_i = n_t + i;
        term = term * four * n_plus_i * n_minus_i / (two_i_plus_1 * two_i_plus_2);
        current_inner_sum += term;
        d_coeffs[k] = n_t * current_inner_sum;
    }
    double d_n = d_coeffs[N];
    double sum = 0.0, compensation = 0.0;
    for (int k = 0; k < N; ++k) {
        double k_plus_1 = (double)(k + 1);
        double sign = (k % 2 == 0) ? one : -one;
        double current_term = sign * (d_coeffs[k] - d_n) / pow(k_plus_1, s);
        double y = current_term - compensation;
        double t = sum + y;
        compensation = (t - sum) - y;
        sum = t;
    }
    return -sum / (d_n * denom);
}

/* ---- zeta.rs:15-89 ---- */
double saf_o_zeta(double x) {
    const double one = 1.0;
    const double threshold = 1e-10;
    double delta = x - one;
    if (fabs(delta) < threshold) return signed_infinity_from_delta(delta);

    if (x < 0.0) {
        const double pi = SAF_PI, two = 2.0;
        double gs = saf_o_gamma(one - x);
        double z = saf_o_zeta(one - x);
        double term1 = pow(two, x);
        double term2 = pow(pi, x - one);
        double term3 = sin(pi * x / two);
        return term1 * term2 * term3 * gs * z;
    }
    if (x > one && x <= 1.5) {
        const int n_terms = 100;
        double sum = 0.0, compensation = 0.0;
        for (int k = 1; k <= n_terms; ++k) {
            double k_t = (double)k;
            double term = one / pow(k_t, x);
            double y = term - compensation;
            double t = sum + y;
            compensation = (t - sum) - y;
            sum = t;
        }
        double n = (double)n_terms;
        double n_pow_x = pow(n, x);
        double n_pow_1_minus_x = pow(n, one - x);
        double em_integral = n_pow_1_minus_x / (x - one);
        double em_boundary = 0.5 / n_pow_x;
        double b2_correction = x / (12.0 * pow(n, x + one));
        double s_plus_1 = x + one;
        double s_plus_2 = x + 2.0;
        double b4_correction = -x * s_plus_1 * s_plus_2 / (720.0 * pow(n, x + 3.0));
        return sum + em_integral + em_boundary + b2_correction + b4_correction;
    }
    return zeta_borwein(x);
}

/* ---- zeta.rs:308-326 ---- */
static double zeta_reflection_base(double s) {
    const double pi = SAF_PI, two = 2.0, half = 0.5, one = 1.0;
    double one_minus_s = one - s;
    double a_term = pow(two, s);
    double b_term = pow(pi, s - one);
    double c_term = sin(pi * s * half);
    double d_term = saf_o_gamma(one_minus_s);
    double zeta_term = saf_o_zeta(one_minus_s);
    return a_term * b_term * c_term * d_term * zeta_term;
}

/* ---- zeta.rs:331-374 ---- */
static double centered_finite_diff(int32_t n, double s, double h) {
    const double two = 2.0, four = 4.0;
    if (n == 1) {
        double zp = zeta_reflection_base(s + h);
        double zm = zeta_reflection_base(s - h);
        return (zp - zm) / (two * h);
    }
    if (n == 2) {
        double zp = zeta_reflection_base(s + h);
        double zc = zeta_reflection_base(s);
        double zm = zeta_reflection_base(s - h);
        return (zp - two * zc + zm) / (h * h);
    }
    if (n == 3) {
        double zp2 = zeta_reflection_base(s + two * h);
        double zp1 = zeta_reflection_base(s + h);
        double zm1 = zeta_reflection_base(s - h);
        double zm2 = zeta_reflection_base(s - two * h);
        return (-zp2 + two * zp1 - two * zm1 + zm2) / (two * h * h * h);
    }
    double two_h = two * h;
    double zp2 = zeta_reflection_base(s + two_h);
    double zp1 = zeta_reflection_base(s + h);
    double zc = zeta_reflection_base(s);
    double zm1 = zeta_reflection_base(s - h);
    double zm2 = zeta_reflection_base(s - two_h);
    if (n == 4) {
        const double six = 6.0;
        return (zp2 - four * zp1 + six * zc - four * zm1 + zm2) / (h * h * h * h);
    }
    double dp = centered_finite_diff(n - 1, s + h, h);
    double dm = centered_finite_diff(n - 1, s - h, h);
    return (dp - dm) / (two * h);
}

/* ---- zeta.rs:276-302 ---- */
static double zeta_deriv_reflection(int32_t n, double s) {
    const double one = 1.0, two = 2.0, four = 4.0;
    const double h = 1e-7;
    if (n == 1) {
        double zp = zeta_reflection_base(s + h);
        double zm = zeta_reflection_base(s - h);
        return (zp - zm) / (two * h);
    } else if (n == 2) {
        double zp = zeta_reflection_base(s + h);
        double zc = zeta_reflection_base(s);
        double zm = zeta_reflection_base(s - h);
        return (zp - two * zc + zm) / (h * h);
    }
    double d_h = centered_finite_diff(n, s, h);
    double d_h2 = centered_finite_diff(n, s, h / two);
    double extrapolation = saf_o_powi(four, n) - one;
    return (saf_o_powi(four, n) * d_h2 - d_h) / extrapolation;
}

/* ---- zeta.rs:186-266 ---- */
double saf_o_zeta_deriv(int32_t n, double x) {
    if (n < 0) return NAN;
    if (n == 0) return saf_o_zeta(x);
    if (n > 16) return NAN; /* work cap: the reference needs 2^(n-4) zeta evaluations below s = 1 */
    const double one = 1.0, epsilon = 1e-10;
    double delta = x - one;
    if (fabs(delta) < epsilon) {
        double sign = (n % 2 == 0) ? sign_from_delta(delta) : -1.0;
        return signed_infinity(sign);
    }
    if (x > one) {
        double sum = 0.0, compensation = 0.0;
        const int max_terms = 200;
        for (int k = 1; k <= max_terms; ++k) {
            double k_t = (double)k;
            double ln_k = log(k_t);
            double ln_k_power;
            if (n <= 5) {
                double result = one;
                for (int32_t j = 0; j < n; ++j) result *= ln_k;
                ln_k_power = result;
            } else {
                ln_k_power = pow(ln_k, (double)n);
            }
            double term = ln_k_power / pow(k_t, x);
            double y = term - compensation;
            double t = sum + y;
            compensation = (t - sum) - y;
            sum = t;
            if (k > 50 && fabs(term) < epsilon * 0.01) break;
        }
        double sign = (n % 2 == 0) ? one : -one;
        return sign * sum;
    }
    return zeta_deriv_reflection(n, x);
}

/* ---- lambert_w.rs:13-91 ---- */
double saf_o_lambert_w(double x) {
    const double one = 1.0, e = SAF_E;
    double e_inv = one / e;
    if (x < -e_inv) return NAN;
    if (x == 0.0) return 0.0;
    const double threshold = 1e-12;
    if (fabs(x + e_inv) < threshold) return -one;

    double w;
    if (x < -0.3) {
        const double two = 2.0;
        double arg = fmax(two * (e * x + one), 0.0);
        double p = sqrt(arg);
        const double third = 3.0;
        const double c1 = 11.0 / 72.0;
        w = -one + p - p * p / third + c1 * p * p * p;
    } else if (x < 0.0) {
        const double two = 2.0;
        double p = sqrt(two * (e * x + one));
        w = -one + p;
    } else if (x < one) {
        const double one_point_five = 1.5;
        w = x * (one - x * (one - x * one_point_five));
    } else if (x < 3.0) {
        double l = log(x);
        double l_ln = log(l);
        double safe_l_ln = (l_ln > 0.0) ? l_ln : 0.0;
        w = l - safe_l_ln;
    } else {
        double l1 = log(x);
        double l2 = log(l1);
        w = l1 - l2 + l2 / l1;
    }
    const double tolerance = 1e-15, neg_one = -one, two = 2.0, half = 0.5;
    for (int it = 0; it < 50; ++it) {
        if (w <= neg_one) w = -0.99;
        double ew = exp(w);
        double wew = w * ew;
        double f = wew - x;
        double w1 = w + one;
        if (fabs(w1) < tolerance) break;
        double fp = ew * w1;
        double fpp = ew * (w + two);
        double d = f * fp / (fp * fp - half * f * fpp);
        w -= d;
        if (fabs(d) < tolerance * (one + fabs(w))) break;
    }
    return w;
}

/* ---- elliptic.rs:11-40 ---- */
double saf_o_elliptic_k(double k) {
    const double one = 1.0;
    double k_abs = fabs(k);
    if (isnan(k_abs) || k_abs > one) return NAN;
    if (k_abs == one) return INFINITY;
    double a = one;
    double b = sqrt(one - k * k);
    const double two = 2.0, tolerance = SAF_EPSILON;
    for (int it = 0; it < 25; ++it) {
        double an = (a + b) / two;
        double bn = sqrt(a * b);
        a = an;
        b = bn;
        if (fabs(a - b) < tolerance) break;
    }
    return SAF_PI / (two * a);
}

/* ---- elliptic.rs:48-78 ---- */
double saf_o_elliptic_e(double k) {
    const double one = 1.0;
    double k_abs = fabs(k);
    if (isnan(k_abs) || k_abs > one) return NAN;
    double a = one;
    double b = sqrt(one - k * k);
    double k2 = k * k;
    double sum = one - k2 / 2.0;
    double pow2 = 0.5;
    const double two = 2.0, tolerance = SAF_EPSILON;
    for (int it = 0; it < 25; ++it) {
        double an = (a + b) / two;
        double bn = sqrt(a * b);
        double cn = (a - b) / two;
        sum -= pow2 * cn * cn;
        a = an;
        b = bn;
        pow2 *= two;
        if (fabs(cn) < tolerance) break;
    }
    return SAF_PI / (two * a) * sum;
}

/* ---- bessel.rs:690-714 ---- */
static double poly_horner(double x, const double *c, int n) {
    double sum = 0.0;
    for (int i = n - 1; i >= 0; --i) sum = sum * x + c[i];
    return sum;
}
static double rational_poly(double x, const double *num, int nn, const double *den, int nd) {
    double n = poly_horner(x, num, nn);
    double d = poly_horner(x, den, nd);
    return n / d;
}

/* ---- bessel.rs:152-209 ---- */
static double bessel_j0(double x) {
    static const double SMALL_NUM[6] = {57568490574.0, -13362590354.0, 651619640.7,
                                        -11214424.18,  77392.33017,    -184.9052456};
    static const double SMALL_DEN[6] = {57568490411.0, 1029532985.0, 9494680.718,
                                        59272.64853,   267.8532712,  1.0};
    static const double LARGE_P_COS[5] = {1.0, -0.1098628627e-2, 0.2734510407e-4,
                                          -0.2073370639e-5, 0.2093887211e-6};
    static const double LARGE_P_SIN[5] = {-0.1562499995e-1, 0.1430488765e-3, -0.6911147651e-5,
                                          0.7621095161e-6, 0.934935152e-7};
    double ax = fabs(x);
    const double eight = 8.0;
    if (ax < eight) {
        double y = x * x;
        return rational_poly(y, SMALL_NUM, 6, SMALL_DEN, 6);
    }
    double z = eight / ax;
    double y = z * z;
    double shift = 0.785398164;
    double xx = ax - shift;
    double term_sqrt = sqrt(SAF_FRAC_2_PI / ax);
    double p_cos = poly_horner(y, LARGE_P_COS, 5);
    double p_sin = poly_horner(y, LARGE_P_SIN, 5);
    return term_sqrt * (cos(xx) * p_cos - z * sin(xx) * p_sin);
}

/* ---- bessel.rs:216-274 ---- */
static double bessel_j1(double x) {
    static const double SMALL_NUM[6] = {72362614232.0, -7895059235.0, 242396853.1,
                                        -2972611.439,  15704.48260,   -30.16036606};
    static const double SMALL_DEN[6] = {144725228442.0, 2300535178.0, 18583304.74,
                                        99447.43394,    376.9991397,  1.0};
    static const double LARGE_P_COS[5] = {1.0, 0.183105e-2, -0.3516396496e-4, 0.2457520174e-5,
                                          -0.240337019e-6};
    static const double LARGE_P_SIN[5] = {0.04687499995, -0.2002690873e-3, 0.8449199096e-5,
                                          -0.88228987e-6, 0.105787412e-6};
    double ax = fabs(x);
    const double eight = 8.0;
    if (ax < eight) {
        double y = x * x;
        double term = rational_poly(y, SMALL_NUM, 6, SMALL_DEN, 6);
        return x * term;
    }
    double z = eight / ax;
    double y = z * z;
    double shift = 2.356194491;
    double xx = ax - shift;
    double term_sqrt = sqrt(SAF_FRAC_2_PI / ax);
    double p_cos = poly_horner(y, LARGE_P_COS, 5);
    double p_sin = poly_horner(y, LARGE_P_SIN, 5);
    double ans = term_sqrt * (cos(xx) * p_cos - z * sin(xx) * p_sin);
    return (x < 0.0) ? -ans : ans;
}

/* ---- bessel.rs:38-65 ---- */
static double bessel_j_forward(int32_t n, double x) {
    int32_t n_abs = n < 0 ? -n : n;
    double j0 = bessel_j0(x);
    if (n_abs == 0) return j0;
    double j1 = bessel_j1(x);
    if (n_abs == 1) return (n < 0) ? -j1 : j1;
    double jp = j0, jc = j1;
    const double two = 2.0;
    double k_t = 1.0;
    for (int32_t i = 1; i < n_abs; ++i) {
        double jn = (two * k_t / x) * jc - jp;
        jp = jc;
        jc = jn;
        k_t += 1.0;
    }
    return (n < 0 && n_abs % 2 == 1) ? -jc : jc;
}

/* ---- bessel.rs:128-141: Rust `as i32` saturates ---- */
static int32_t sat_i32(double v) {
    if (isnan(v)) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}
static int32_t compute_miller_start(int32_t n) {
    double n_f = (double)n;
    const double c = 50.0;
    const int32_t m = 15;
    int32_t sqrt_round = sat_i32(round(sqrt(c * n_f)));
    return n + sqrt_round + m;
}

/* ---- bessel.rs:68-125 ---- */
static double bessel_j_miller(int32_t n, double x) {
    int32_t n_abs = n < 0 ? -n : n;
    const double two = 2.0;
    int32_t n_start = compute_miller_start(n_abs);
    double j_next = 0.0, j_curr = 1e-30;
    double result = 0.0, sum = 0.0, compensation = 0.0;
    double k_t = (double)n_start;
    for (int32_t k = n_start; k >= 0; --k) {
        double j_prev = (two * k_t / x) * j_curr - j_next;
        if (k == n_abs) result = j_curr;
        if (k == 0) {
            double y = j_curr - compensation;
            double t = sum + y;
            compensation = (t - sum) - y;
            sum = t;
        } else if (k % 2 == 0) {
            double term = two * j_curr;
            double y = term - compensation;
            double t = sum + y;
            compensation = (t - sum) - y;
            sum = t;
        }
        j_next = j_curr;
        j_curr = j_prev;
        k_t -= 1.0;
    }
    double scale = 1.0 / sum;
    double normalized_result = result * scale;
    return (n < 0 && n_abs % 2 == 1) ? -normalized_result : normalized_result;
}

/* ---- bessel.rs:16-35 ---- */
double saf_o_bessel_j(int32_t n, double x) {
    if (n > SAF_MAX_ORDER || n < -SAF_MAX_ORDER) return NAN; /* work cap */
    int32_t n_abs = n < 0 ? -n : n;
    double ax = fabs(x);
    const double threshold = 1e-10;
    if (ax < threshold) return (n_abs == 0) ? 1.0 : 0.0;
    if ((double)n_abs <= ax) return bessel_j_forward(n, x);
    return bessel_j_miller(n, x);
}

/* ---- bessel.rs:286-339 ---- */
static double bessel_y0(double x) {
    static const double SMALL_NUM[6] = {-2957821389.0, 7062834065.0, -512359803.6,
                                        10879881.29,   -86327.92757, 228.4622733};
    static const double SMALL_DEN[6] = {40076544269.0, 745249964.8, 7189466.438,
                                        47447.26470,   226.1030244, 1.0};
    static const double LARGE_P_SIN[5] = {1.0, -0.1098628627e-2, 0.2734510407e-4,
                                          -0.2073370639e-5, 0.2093887211e-6};
    static const double LARGE_P_COS[5] = {-0.1562499995e-1, 0.1430488765e-3, -0.6911147651e-5,
                                          0.7621095161e-6, 0.934935152e-7};
    const double eight = 8.0;
    if (x < eight) {
        double y = x * x;
        double term = rational_poly(y, SMALL_NUM, 6, SMALL_DEN, 6);
        return term + SAF_FRAC_2_PI * bessel_j0(x) * log(x);
    }
    double z = eight / x;
    double y = z * z;
    double shift = 0.785398164;
    double xx = x - shift;
    double term_sqrt = sqrt(SAF_FRAC_2_PI / x);
    double p_sin = poly_horner(y, LARGE_P_SIN, 5);
    double p_cos = poly_horner(y, LARGE_P_COS, 5);
    return term_sqrt * (sin(xx) * p_sin + z * cos(xx) * p_cos);
}

/* ---- bessel.rs:344-400 ---- */
static double bessel_y1(double x) {
    static const double SMALL_NUM[6] = {-0.4900604943e13, 0.1275274390e13, -0.5153438139e11,
                                        0.7349264551e9,   -0.4237922726e7, 0.8511937935e4};
    static const double SMALL_DEN[7] = {0.2499580570e14, 0.4244419664e12, 0.3733650367e10,
                                        0.2245904002e8,  0.1020426050e6,  0.3549632885e3, 1.0};
    static const double LARGE_P_SIN[5] = {1.0, 0.183105e-2, -0.3516396496e-4, 0.2457520174e-5,
                                          -0.240337019e-6};
    static const double LARGE_P_COS[5] = {0.04687499995, -0.2002690873e-3, 0.8449199096e-5,
                                          -0.88228987e-6, 0.105787412e-6};
    const double eight = 8.0;
    if (x < eight) {
        double y = x * x;
        double term_poly = rational_poly(y, SMALL_NUM, 6, SMALL_DEN, 7);
        double term = x * term_poly;
        return term + SAF_FRAC_2_PI * (bessel_j1(x) * log(x) - 1.0 / x);
    }
    double z = eight / x;
    double y = z * z;
    double shift = 2.356194491;
    double xx = x - shift;
    double term_sqrt = sqrt(SAF_FRAC_2_PI / x);
    double p_sin = poly_horner(y, LARGE_P_SIN, 5);
    double p_cos = poly_horner(y, LARGE_P_COS, 5);
    return term_sqrt * (sin(xx) * p_sin + z * cos(xx) * p_cos);
}

/* ---- bessel.rs:256-281 ---- */
double saf_o_bessel_y(int32_t n, double x) {
    if (n > SAF_MAX_ORDER || n < -SAF_MAX_ORDER) return NAN; /* work cap */
    if (x <= 0.0) return NAN;
    int32_t n_abs = n < 0 ? -n : n;
    double y0 = bessel_y0(x);
    if (n_abs == 0) return y0;
    double y1 = bessel_y1(x);
    if (n_abs == 1) return (n < 0) ? -y1 : y1;
    double yp = y0, yc = y1;
    const double two = 2.0;
    double k_t = 1.0;
    for (int32_t i = 1; i < n_abs; ++i) {
        double yn = (two * k_t / x) * yc - yp;
        yp = yc;
        yc = yn;
        k_t += 1.0;
    }
    return (n < 0 && n_abs % 2 == 1) ? -yc : yc;
}

/* ---- bessel.rs:482-518 ---- */
static double bessel_i0(double x) {
    static const double SMALL[7] = {1.0,       3.5156229, 3.0899424, 1.2067492,
                                    0.2659732, 0.0360768, 0.0045813};
    static const double LARGE[9] = {0.39894228,  0.01328592, 0.00225319,  -0.00157565, 0.00916281,
                                    -0.02057706, 0.02635537, -0.01647633, 0.00392377};
    double ax = fabs(x);
    const double three_seven_five = 3.75;
    if (ax < three_seven_five) {
        double y = saf_o_powi(x / three_seven_five, 2);
        return poly_horner(y, SMALL, 7);
    }
    double y = three_seven_five / ax;
    double term = exp(ax) / sqrt(ax);
    return term * poly_horner(y, LARGE, 9);
}

/* ---- bessel.rs:523-560 ---- */
static double bessel_i1(double x) {
    static const double SMALL[7] = {0.5,        0.87890594, 0.51498869, 0.15084934,
                                    0.02658733, 0.00301532, 0.00032411};
    static const double LARGE[9] = {0.39894228, -0.03988024, -0.00362018, 0.00163801, -0.01031555,
                                    0.02282967, -0.02895312, 0.01787654,  -0.00420059};
    double ax = fabs(x);
    const double three_seven_five = 3.75;
    double ans;
    if (ax < three_seven_five) {
        double y = saf_o_powi(x / three_seven_five, 2);
        ans = ax * poly_horner(y, SMALL, 7);
    } else {
        double y = three_seven_five / ax;
        double term = exp(ax) / sqrt(ax);
        ans = term * poly_horner(y, LARGE, 9);
    }
    return (x < 0.0) ? -ans : ans;
}

/* ---- bessel.rs:410-477 ---- */
double saf_o_bessel_i(int32_t n, double x) {
    if (n > SAF_MAX_ORDER || n < -SAF_MAX_ORDER) return NAN; /* work cap */
    int32_t n_abs = n < 0 ? -n : n;
    if (n_abs == 0) return bessel_i0(x);
    if (n_abs == 1) return bessel_i1(x);
    const double threshold = 1e-10;
    if (fabs(x) < threshold) return 0.0;
    const double two = 2.0;
    int32_t sqrt_term = sat_i32(sqrt((double)(40 * n_abs)));
    int32_t n_start = n_abs + sqrt_term + 10;
    if (n_start < n_abs + 20) n_start = n_abs + 20;
    double i_next = 0.0, i_curr = 1e-30, result = 0.0, sum = 0.0;
    double k_t = (double)n_start;
    for (int32_t k = n_start; k >= 0; --k) {
        double i_prev = (two * k_t / x) * i_curr + i_next;
        if (k == n_abs) result = i_curr;
        if (k == 0) {
            sum += i_curr;
        } else if (k % 2 == 0) {
            sum += two * i_curr;
        }
        i_next = i_curr;
        i_curr = i_prev;
        k_t -= 1.0;
    }
    double i0_actual = bessel_i0(x);
    double scale = i0_actual / sum;
    return result * scale;
}

/* ---- bessel.rs:595-636 ---- */
static double bessel_k0(double x) {
    static const double COEFFS[7] = {-0.57721566, 0.42278420, 0.23069756, 0.03488590,
                                     0.00262698,  0.00010750, 0.0000074};
    static const double COEFFS_LARGE[8] = {1.25331414, -0.07832358, 0.02189568, -0.01062446,
                                           0.00587872, -0.00251540, 0.00053208, -0.000025200};
    const double two = 2.0;
    if (x <= two) {
        const double four = 4.0;
        double y = x * x / four;
        double i0 = bessel_i0(x);
        double ln_term = -log(x / two) * i0;
        double poly = poly_horner(y, COEFFS, 7);
        return ln_term + poly;
    }
    double y = two / x;
    double term = exp(-x) / sqrt(x);
    return term * poly_horner(y, COEFFS_LARGE, 8);
}

/* ---- bessel.rs:639-685 ---- */
static double bessel_k1(double x) {
    static const double COEFFS[7] = {1.0,         0.15443144,  -0.67278579, -0.18156897,
                                     -0.01919402, -0.00110404, -0.00004686};
    static const double COEFFS_LARGE[8] = {1.25331414,  0.23498619, -0.03655620, 0.01504268,
                                           -0.00780353, 0.00325614, -0.00068245, 0.0000316};
    const double two = 2.0;
    if (x <= two) {
        const double four = 4.0;
        double y = x * x / four;
        double term1 = log(x) * bessel_i1(x);
        double term2 = 1.0 / x;
        double poly = poly_horner(y, COEFFS, 7);
        return term1 + term2 * poly;
    }
    double y = two / x;
    double term = exp(-x) / sqrt(x);
    return term * poly_horner(y, COEFFS_LARGE, 8);
}

/* ---- bessel.rs:565-590 ---- */
double saf_o_bessel_k(int32_t n, double x) {
    if (n > SAF_MAX_ORDER || n < -SAF_MAX_ORDER) return NAN; /* work cap */
    if (x <= 0.0) return NAN;
    int32_t n_abs = n < 0 ? -n : n;
    double k0 = bessel_k0(x);
    if (n_abs == 0) return k0;
    double k1 = bessel_k1(x);
    if (n_abs == 1) return k1;
    double kp = k0, kc = k1;
    const double two = 2.0;
    double k_t = 1.0;
    for (int32_t i = 1; i < n_abs; ++i) {
        double kn = kp + (two * k_t / x) * kc;
        kp = kc;
        kc = kn;
        k_t += 1.0;
    }
    return kc;
}

/* ---- polynomials.rs:10-31 ---- */
double saf_o_hermite(int32_t n, double x) {
    if (n > SAF_MAX_ORDER) return NAN; /* work cap */
    if (n < 0) return NAN;
    if (n == 0) return 1.0;
    const double two = 2.0;
    double term1 = two * x;
    if (n == 1) return term1;
    double h0 = 1.0, h1 = term1;
    for (int32_t k = 1; k < n; ++k) {
        double f_k = (double)k;
        double h2 = (two * x * h1) - (two * f_k * h0);
        h0 = h1;
        h1 = h2;
    }
    return h1;
}

/* ---- polynomials.rs:39-88 ---- */
double saf_o_assoc_legendre(int32_t l, int32_t m, double x) {
    if (l > SAF_MAX_ORDER) return NAN; /* work cap */
    int32_t m_abs = m < 0 ? -m : m;
    if (l < 0 || m_abs > l) return NAN;
    double x_abs = fabs(x);
    if (isnan(x_abs) || x_abs > 1.0) return NAN;
    double pmm = 1.0;
    const double one = 1.0;
    if (m_abs > 0) {
        double sqx = sqrt(one - x * x);
        double fact = 1.0;
        const double two = 2.0;
        for (int32_t i = 1; i <= m_abs; ++i) {
            pmm = pmm * (-fact) * sqx;
            fact += two;
        }
    }
    if (l == m_abs) return pmm;
    double two_m_plus_1 = (double)(2 * m_abs + 1);
    double pmmp1 = x * two_m_plus_1 * pmm;
    if (l == m_abs + 1) return pmmp1;
    double pll = 0.0, pmm_prev = pmm, pmm_curr = pmmp1;
    for (int32_t ll = m_abs + 2; ll <= l; ++ll) {
        double f_ll = (double)ll;
        double f_m_abs = (double)m_abs;
        double term1_fact = (double)(2 * ll - 1);
        double term2_fact = (double)(ll + m_abs - 1);
        double denom = f_ll - f_m_abs;
        pll = (x * term1_fact * pmm_curr - term2_fact * pmm_prev) / denom;
        pmm_prev = pmm_curr;
        pmm_curr = pll;
    }
    return pll;
}

/* ---- polynomials.rs:96-128 ---- */
double saf_o_spherical_harmonic(int32_t l, int32_t m, double theta, double phi) {
    if (l > SAF_MAX_ORDER) return NAN; /* work cap */
    int32_t m_abs = m < 0 ? -m : m;
    if (l < 0 || m_abs > l) return NAN;
    double cos_theta = cos(theta);
    if (isnan(cos_theta)) return NAN;
    double plm = saf_o_assoc_legendre(l, m, cos_theta);
    double fact_lm = 1.0;
    for (int32_t i = 1; i <= l - m_abs; ++i) fact_lm *= (double)i;
    double fact_lplusm = 1.0;
    for (int32_t i = 1; i <= l + m_abs; ++i) fact_lplusm *= (double)i;
    const double four = 4.0;
    double two_l_plus_1 = (double)(2 * l + 1);
    double norm_sq = (two_l_plus_1 / (four * SAF_PI)) * (fact_lm / fact_lplusm);
    double norm = sqrt(norm_sq);
    double m_phi = (double)m * phi;
    return norm * plm * cos(m_phi);
}
