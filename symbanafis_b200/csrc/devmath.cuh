// devmath.cuh -- device-side leaf math of the interpreter kernel.
//
// Semantics follow the reference evaluator, not a naive libm mapping:
//   scalar arms of dispatch_loop!      src/evaluator/logic/bytecode/execute/engine/macros.rs:359-411
//   eval_builtin1..4                   .../execute/engine/builtins.rs:38-139
//   eval_sinc, round_to_i32            .../execute/engine/helpers.rs:9-25
//   special functions (T = f64)        src/math/logic/functions/{erf,gamma,beta,polygamma,zeta,
//                                      lambert_w,elliptic,bessel,polynomials,polar}.rs
// Elementary transcendentals come from the CUDA math library (<= 2 ulp, documented), the special
// functions are re-implemented here from the reference's algorithms so that their (sometimes
// peculiar) numerical behaviour -- e.g. the 30-term Taylor erf that diverges beyond |x| ~ 3 -- is
// reproduced.  The translation unit is compiled with -fmad=false: rustc never contracts a*b+c, and
// only the MulAdd-family opcodes may fuse.
#pragma once
#include <cfloat>
#include <cstdint>

#include "saf_isa.h"

namespace saf {
namespace dm {

#define SAF_DEV __device__ __forceinline__
#define SAF_DEVFN __device__ __noinline__

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kE = 2.71828182845904523536028747135266250;
constexpr double kFracPi2 = 1.57079632679489661923132169163975144;
constexpr double kFrac2Pi = 0.636619772367581343075535053490057448;
constexpr double kEps = 1e-14; // crate::EPSILON, src/lib.rs:321

SAF_DEV double nan_() { return __longlong_as_double(0x7ff8000000000000LL); }
SAF_DEV double inf_() { return __longlong_as_double(0x7ff0000000000000LL); }

// f64::powi -> compiler-rt __powidf2: square-and-multiply from the low bit, reciprocal last
SAF_DEV double powi(double a, int b) {
    const bool recip = b < 0;
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

SAF_DEV double signum(double x) { return isnan(x) ? x : copysign(1.0, x); }
SAF_DEV double fract(double x) { return x - trunc(x); }

// Rust std f64::asinh / acosh / atanh formulas
SAF_DEV double rs_asinh(double x) {
    double ax = fabs(x);
    double ix = 1.0 / ax;
    return copysign(log1p(ax + (ax / (hypot(1.0, ix) + ix))), x);
}
SAF_DEV double rs_acosh(double x) {
    if (x < 1.0) return nan_();
    return log(x + (sqrt(x - 1.0) * sqrt(x + 1.0)));
}
SAF_DEV double rs_atanh(double x) { return 0.5 * log1p((2.0 * x) / (1.0 - x)); }

SAF_DEV double sinc(double x) { return fabs(x) < kEps ? 1.0 : sin(x) / x; }

SAF_DEV bool round_to_i32(double x, int &out) {
    double r = round(x);
    if (!(r >= -2147483648.0 && r <= 2147483647.0)) return false;
    out = (int)r;
    return true;
}

// ---- math/logic/functions/helpers.rs ----
SAF_DEV double sign_from_delta(double d) { return isnan(d) ? nan_() : (signbit(d) ? -1.0 : 1.0); }
SAF_DEV double signed_infinity(double s) { return isnan(s) ? nan_() : (signbit(s) ? -inf_() : inf_()); }
SAF_DEV long long to_i64_or0(double v) {
    if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return 0;
    return (long long)v;
}
SAF_DEV double gamma_pole_sign(double x) { return (to_i64_or0(-x) & 1) == 0 ? 1.0 : -1.0; }
SAF_DEV double gamma_sign(double x) {
    if (x > 0.0) return 1.0;
    return (to_i64_or0(-x) & 1) == 0 ? -1.0 : 1.0;
}

// ---- erf.rs:9-51: alternating Taylor series, Kahan-compensated, at most 30 terms ----
SAF_DEVFN double erf_ref(double x0) {
    const double sign = signum(x0);
    const double x = fabs(x0);
    const double coeff = 2.0 / sqrt(kPi);
    const double xx = x * x;
    double sum = 0.0, comp = 0.0, fact = 1.0, power = x;
    for (int n = 0; n < 30; ++n) {
        double term = power / (fact * (double)(2 * n + 1));
        if (isnan(term) || isinf(term)) break;
        double st = (n & 1) ? -term : term;
        double y = st - comp;
        double t = sum + y;
        comp = (t - sum) - y;
        sum = t;
        fact *= (double)(n + 1);
        power *= xx;
        if (fabs(term) < DBL_EPSILON) break;
    }
    return sign * coeff * sum;
}

// ---- erf.rs:64-126 ----
SAF_DEVFN double erfc_ref(double x) {
    if (isnan(x)) return x;
    if (isinf(x)) return signbit(x) ? 2.0 : 0.0;
    const double ax = fabs(x);
    if (x > 28.0) return 0.0;
    if (ax < 1.5) return 1.0 - erf_ref(x);
    // continued fraction (modified Lentz) on |x| >= 1.5; the reference recurses once for x < 0
    const double tiny = 1e-30;
    double f = ax;
    if (fabs(f) < tiny) f = tiny;
    double c = f, d = 0.0;
    for (int j = 1; j < 200; ++j) {
        double a = 0.5 * (double)j;
        d = ax + a * d;
        if (fabs(d) < tiny) d = tiny;
        d = 1.0 / d;
        c = ax + a / c;
        if (fabs(c) < tiny) c = tiny;
        double delta = c * d;
        f *= delta;
        if (fabs(delta - 1.0) < DBL_EPSILON) break;
    }
    double pos = (exp(-ax * ax) / sqrt(kPi)) / f;
    return x < 0.0 ? 2.0 - pos : pos;
}

// ---- gamma.rs ----
SAF_DEV double lanczos_ag(double xm1) {
    const double c[9] = {0.9999999999998099,   676.5203681218851,    -1259.1392167224028,
                         771.3234287776531,    -176.6150291621406,   12.507343278686905,
                         -0.13857109526572012, 9.984369578019572e-6, 1.5056327351493116e-7};
    double ag = c[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) ag += c[i] / (xm1 + (double)i);
    return ag;
}
SAF_DEV double gamma_lanczos(double x) { // x >= 0.5, gamma.rs:57-64
    double xm1 = x - 1.0;
    double ag = lanczos_ag(xm1);
    double t = xm1 + 7.0 + 0.5;
    return sqrt(2.0 * kPi) * pow(t, xm1 + 0.5) * exp(-t) * ag;
}
SAF_DEVFN double gamma_ref(double x) { // gamma.rs:33-65
    if (x <= 0.0 && fract(x) == 0.0) {
        double delta = x - round(x);
        return signed_infinity(gamma_pole_sign(x) * sign_from_delta(delta));
    }
    if (x < 0.5) {
        // reflection; 1 - x > 0.5 so the recursion of the reference has depth one -- except that
        // 1 - x can round to a non-positive integer pole check, which the inner call repeats
        double y = 1.0 - x;
        double gy;
        if (y <= 0.0 && fract(y) == 0.0) {
            double delta = y - round(y);
            gy = signed_infinity(gamma_pole_sign(y) * sign_from_delta(delta));
        } else if (y < 0.5) {
            gy = nan_(); // unreachable: x < 0.5 implies 1 - x >= 0.5
        } else {
            gy = gamma_lanczos(y);
        }
        return kPi / (sin(kPi * x) * gy);
    }
    return gamma_lanczos(x);
}
SAF_DEV double lgamma_pos(double x) { // x >= 0.5, gamma.rs:94-102
    double xm1 = x - 1.0;
    double ag = lanczos_ag(xm1);
    double t = xm1 + 7.0 + 0.5;
    return log(sqrt(2.0 * kPi)) + (xm1 + 0.5) * log(t) - t + log(fabs(ag));
}
SAF_DEVFN double lgamma_ref(double x) { // gamma.rs:84-103
    if (x <= 0.0 && fract(x) == 0.0) return inf_();
    if (x < 0.5) {
        double sin_term = fabs(sin(kPi * x));
        double y = 1.0 - x;
        double ly = (y <= 0.0 && fract(y) == 0.0) ? inf_() : lgamma_pos(y);
        return log(kPi) - log(sin_term) - ly;
    }
    return lgamma_pos(x);
}

// ---- beta.rs:11-33 ----
SAF_DEVFN double beta_ref(double x, double y) {
    if (x <= 0.0 && fract(x) == 0.0) return signed_infinity(gamma_sign(x));
    if (y <= 0.0 && fract(y) == 0.0) return signed_infinity(gamma_sign(y));
    double xy = x + y;
    if (xy <= 0.0 && fract(xy) == 0.0) return 0.0;
    double sign = gamma_sign(x) * gamma_sign(y) * gamma_sign(xy);
    double lbeta = lgamma_ref(x) + lgamma_ref(y) - lgamma_ref(xy);
    return sign * exp(lbeta);
}

// ---- polygamma.rs ----
SAF_DEV double digamma_pos(double x) { // x >= 0.5, polygamma.rs:25-40
    double xv = x, result = 0.0;
    while (xv < 6.0) {
        result -= 1.0 / xv;
        xv += 1.0;
    }
    result += log(xv) - 0.5 / xv;
    double x2 = xv * xv;
    double t1 = 1.0 / (12.0 * x2);
    double t2 = 1.0 / (120.0 * x2 * x2);
    double t3 = 1.0 / (252.0 * x2 * x2 * x2);
    return result - t1 + t2 - t3;
}
SAF_DEVFN double digamma_ref(double x) { // polygamma.rs:11-41
    if (x <= 0.0 && fract(x) == 0.0) {
        double delta = x - round(x);
        return signed_infinity(-sign_from_delta(delta));
    }
    if (x < 0.5) {
        double y = 1.0 - x;
        double dy;
        if (y <= 0.0 && fract(y) == 0.0) {
            double delta = y - round(y);
            dy = signed_infinity(-sign_from_delta(delta));
        } else {
            dy = digamma_pos(y);
        }
        return dy - kPi * cos(kPi * x) / sin(kPi * x);
    }
    return digamma_pos(x);
}
SAF_DEVFN double trigamma_ref(double x) { // polygamma.rs:49-68
    if (x <= 0.0 && fract(x) == 0.0) return inf_();
    double xv = x, r = 0.0;
    while (xv < 6.0) {
        r += 1.0 / (xv * xv);
        xv += 1.0;
    }
    double x2 = xv * xv;
    return r + 1.0 / xv + 0.5 / x2 + 1.0 / (6.0 * x2 * xv) - 1.0 / (30.0 * x2 * x2 * xv) +
           1.0 / (42.0 * x2 * x2 * x2 * xv);
}
SAF_DEVFN double tetragamma_ref(double x) { // polygamma.rs:75-94
    if (x <= 0.0 && fract(x) == 0.0) {
        double delta = x - round(x);
        return signed_infinity(-sign_from_delta(delta));
    }
    double xv = x, r = 0.0;
    while (xv < 6.0) {
        r -= 2.0 / (xv * xv * xv);
        xv += 1.0;
    }
    double x2 = xv * xv;
    return r - 1.0 / x2 + 1.0 / (x2 * xv) + 1.0 / (2.0 * x2 * x2) + 1.0 / (6.0 * x2 * x2 * xv);
}
SAF_DEVFN double polygamma_ref(int n, double x) { // polygamma.rs:102-208
    if (n < 0) return nan_();
    if (n == 0) return digamma_ref(x);
    if (n == 1) return trigamma_ref(x);
    if (x <= 0.0 && fract(x) == 0.0) {
        double delta = x - round(x);
        double sign = (n % 2 == 0) ? -sign_from_delta(delta) : 1.0;
        return signed_infinity(sign);
    }
    double xv = x, r = 0.0;
    const double sign = ((n + 1) % 2 == 0) ? 1.0 : -1.0;
    double factorial = 1.0;
    for (int i = 1; i <= n; ++i) factorial *= (double)i;
    const int np1 = n + 1;
    while (xv < 15.0) {
        r += sign * factorial / powi(xv, np1);
        xv += 1.0;
    }
    const double asym_sign = (n % 2 == 0) ? -1.0 : 1.0;
    const double bnum[5] = {1.0, -1.0, 1.0, -1.0, 5.0};
    const double bden[5] = {6.0, 30.0, 42.0, 30.0, 66.0};
    double nm1_fact = 1.0;
    for (int i = 1; i < n; ++i) nm1_fact *= (double)i;
    double sum = nm1_fact / powi(xv, n);
    sum += factorial / (2.0 * powi(xv, np1));
    double xpow = powi(xv, n + 2);
    double fact_ratio = factorial * (double)(n + 1);
    double prev = DBL_MAX;
    for (int k = 0; k < 5; ++k) {
        int two_k = 2 * (k + 1);
        double f2k = 1.0;
        for (int i = 1; i <= two_k; ++i) f2k *= (double)i;
        double term = (bnum[k] / bden[k]) * fact_ratio / (f2k * xpow);
        if (fabs(term) > prev) break;
        prev = fabs(term);
        sum += term;
        xpow *= xv * xv;
        fact_ratio *= (double)(n + two_k) * (double)(n + two_k + 1);
    }
    return r + asym_sign * sum;
}

// ---- zeta.rs ----
SAF_DEVFN double zeta_borwein(double s) { // zeta.rs:107-163
    const int N = 14;
    double denom = 1.0 - pow(2.0, 1.0 - s);
    if (fabs(denom) < 1e-15) return signed_infinity(sign_from_delta(s - 1.0));
    double d[N + 1];
    const double n_t = (double)N;
    double term = 1.0 / n_t;
    double inner = term;
    d[0] = n_t * inner;
    for (int k = 1; k <= N; ++k) {
        double i = (double)(k - 1);
        term = term * 4.0 * (n_t + i) * (n_t - i) / ((double)(2 * k - 1) * (double)(2 * k));
        inner += term;
        d[k] = n_t * inner;
    }
    const double d_n = d[N];
    double sum = 0.0, comp = 0.0;
    for (int k = 0; k < N; ++k) {
        double sign = (k % 2 == 0) ? 1.0 : -1.0;
        double cur = sign * (d[k] - d_n) / pow((double)(k + 1), s);
        double y = cur - comp;
        double t = sum + y;
        comp = (t - sum) - y;
        sum = t;
    }
    return -sum / (d_n * denom);
}
SAF_DEVFN double zeta_nonneg(double x) { // zeta.rs:38-88, x >= 0 and away from the pole
    if (x > 1.0 && x <= 1.5) {
        double sum = 0.0, comp = 0.0;
        for (int k = 1; k <= 100; ++k) {
            double term = 1.0 / pow((double)k, x);
            double y = term - comp;
            double t = sum + y;
            comp = (t - sum) - y;
            sum = t;
        }
        const double n = 100.0;
        double n_pow_x = pow(n, x);
        double n_pow_1mx = pow(n, 1.0 - x);
        double em_integral = n_pow_1mx / (x - 1.0);
        double em_boundary = 0.5 / n_pow_x;
        double b2 = x / (12.0 * pow(n, x + 1.0));
        double b4 = -x * (x + 1.0) * (x + 2.0) / (720.0 * pow(n, x + 3.0));
        return sum + em_integral + em_boundary + b2 + b4;
    }
    return zeta_borwein(x);
}
SAF_DEVFN double zeta_ref(double x) { // zeta.rs:15-89
    double delta = x - 1.0;
    if (fabs(delta) < 1e-10) return signed_infinity(sign_from_delta(delta));
    if (x < 0.0) {
        double y = 1.0 - x; // > 1: the reference's recursion has depth one
        double gs = gamma_ref(y);
        double dy = y - 1.0;
        double z = (fabs(dy) < 1e-10) ? signed_infinity(sign_from_delta(dy)) : zeta_nonneg(y);
        double term1 = pow(2.0, x);
        double term2 = pow(kPi, x - 1.0);
        double term3 = sin(kPi * x / 2.0);
        return term1 * term2 * term3 * gs * z;
    }
    return zeta_nonneg(x);
}
SAF_DEVFN double zeta_reflection_base(double s) { // zeta.rs:308-326
    double oms = 1.0 - s;
    double a = pow(2.0, s);
    double b = pow(kPi, s - 1.0);
    double c = sin(kPi * s * 0.5);
    double d = gamma_ref(oms);
    double z = zeta_ref(oms);
    return a * b * c * d * z;
}
// zeta.rs:331-374.  The reference recurses for n >= 5 (two calls per level); the device walks the
// same binary recursion tree with an explicit stack of (order, point) frames: the value of a node of
// order n at point s is [node(n-1, s+h) - node(n-1, s-h)] / (2h), leaves have order <= 4.
SAF_DEV double cfd_leaf(int n, double s, double h) {
    if (n == 1) return (zeta_reflection_base(s + h) - zeta_reflection_base(s - h)) / (2.0 * h);
    if (n == 2) {
        double zp = zeta_reflection_base(s + h), zc = zeta_reflection_base(s), zm = zeta_reflection_base(s - h);
        return (zp - 2.0 * zc + zm) / (h * h);
    }
    if (n == 3) {
        double zp2 = zeta_reflection_base(s + 2.0 * h), zp1 = zeta_reflection_base(s + h);
        double zm1 = zeta_reflection_base(s - h), zm2 = zeta_reflection_base(s - 2.0 * h);
        return (-zp2 + 2.0 * zp1 - 2.0 * zm1 + zm2) / (2.0 * h * h * h);
    }
    double two_h = 2.0 * h;
    double zp2 = zeta_reflection_base(s + two_h), zp1 = zeta_reflection_base(s + h);
    double zc = zeta_reflection_base(s), zm1 = zeta_reflection_base(s - h), zm2 = zeta_reflection_base(s - two_h);
    return (zp2 - 4.0 * zp1 + 6.0 * zc - 4.0 * zm1 + zm2) / (h * h * h * h);
}
constexpr int kMaxZetaDerivDepth = 12; // orders up to 16; beyond that the reference needs > 2^12 zeta evaluations
SAF_DEVFN double centered_finite_diff(int n, double s, double h) {
    if (n <= 4) return cfd_leaf(n, s, h);
    int depth = n - 4;
    if (depth > kMaxZetaDerivDepth) return nan_(); // documented deviation: refuse exponential work
    // post-order evaluation of a complete binary tree of `depth` levels
    double val[kMaxZetaDerivDepth + 1];
    unsigned path = 0; // bit i = 1 when level i took the "minus" branch
    int level = 0;
    double cur_s[kMaxZetaDerivDepth + 1];
    cur_s[0] = s;
    unsigned state[kMaxZetaDerivDepth + 1]; // 0 = need plus child, 1 = need minus child, 2 = done
    for (int i = 0; i <= kMaxZetaDerivDepth; ++i) state[i] = 0;
    double plus_val[kMaxZetaDerivDepth + 1];
    (void)path;
    for (;;) {
        if (level == depth) {
            val[level] = cfd_leaf(4, cur_s[level], h);
            --level;
            if (level < 0) break;
            continue;
        }
        if (state[level] == 0) {
            state[level] = 1;
            cur_s[level + 1] = cur_s[level] + h;
            state[level + 1] = 0;
            ++level;
        } else if (state[level] == 1) {
            plus_val[level] = val[level + 1];
            state[level] = 2;
            cur_s[level + 1] = cur_s[level] - h;
            state[level + 1] = 0;
            ++level;
        } else {
            val[level] = (plus_val[level] - val[level + 1]) / (2.0 * h);
            --level;
            if (level < 0) break;
        }
    }
    return val[0];
}
SAF_DEVFN double zeta_deriv_ref(int n, double x) { // zeta.rs:186-302
    if (n < 0) return nan_();
    if (n == 0) return zeta_ref(x);
    double delta = x - 1.0;
    if (fabs(delta) < 1e-10) {
        double sign = (n % 2 == 0) ? sign_from_delta(delta) : -1.0;
        return signed_infinity(sign);
    }
    if (x > 1.0) {
        double sum = 0.0, comp = 0.0;
        for (int k = 1; k <= 200; ++k) {
            double k_t = (double)k;
            double ln_k = log(k_t);
            double lp;
            if (n <= 5) {
                lp = 1.0;
                for (int j = 0; j < n; ++j) lp *= ln_k;
            } else {
                lp = pow(ln_k, (double)n);
            }
            double term = lp / pow(k_t, x);
            double y = term - comp;
            double t = sum + y;
            comp = (t - sum) - y;
            sum = t;
            if (k > 50 && fabs(term) < 1e-10 * 0.01) break;
        }
        return ((n % 2 == 0) ? 1.0 : -1.0) * sum;
    }
    const double h = 1e-7;
    if (n == 1) return (zeta_reflection_base(x + h) - zeta_reflection_base(x - h)) / (2.0 * h);
    if (n == 2) {
        double zp = zeta_reflection_base(x + h), zc = zeta_reflection_base(x), zm = zeta_reflection_base(x - h);
        return (zp - 2.0 * zc + zm) / (h * h);
    }
    double d_h = centered_finite_diff(n, x, h);
    double d_h2 = centered_finite_diff(n, x, h / 2.0);
    double extrapolation = powi(4.0, n) - 1.0;
    return (powi(4.0, n) * d_h2 - d_h) / extrapolation;
}

// ---- lambert_w.rs:13-91 ----
SAF_DEVFN double lambert_w_ref(double x) {
    const double e_inv = 1.0 / kE;
    if (x < -e_inv) return nan_();
    if (x == 0.0) return 0.0;
    if (fabs(x + e_inv) < 1e-12) return -1.0;
    double w;
    if (x < -0.3) {
        double arg = fmax(2.0 * (kE * x + 1.0), 0.0);
        double p = sqrt(arg);
        const double c1 = 11.0 / 72.0;
        w = -1.0 + p - p * p / 3.0 + c1 * p * p * p;
    } else if (x < 0.0) {
        double p = sqrt(2.0 * (kE * x + 1.0));
        w = -1.0 + p;
    } else if (x < 1.0) {
        w = x * (1.0 - x * (1.0 - x * 1.5));
    } else if (x < 3.0) {
        double l = log(x);
        double l_ln = log(l);
        w = l - ((l_ln > 0.0) ? l_ln : 0.0);
    } else {
        double l1 = log(x);
        double l2 = log(l1);
        w = l1 - l2 + l2 / l1;
    }
    const double tol = 1e-15;
    for (int it = 0; it < 50; ++it) {
        if (w <= -1.0) w = -0.99;
        double ew = exp(w);
        double f = w * ew - x;
        double w1 = w + 1.0;
        if (fabs(w1) < tol) break;
        double fp = ew * w1;
        double fpp = ew * (w + 2.0);
        double d = f * fp / (fp * fp - 0.5 * f * fpp);
        w -= d;
        if (fabs(d) < tol * (1.0 + fabs(w))) break;
    }
    return w;
}

// ---- elliptic.rs ----
SAF_DEVFN double elliptic_k_ref(double k) {
    double ka = fabs(k);
    if (isnan(ka) || ka > 1.0) return nan_();
    if (ka == 1.0) return inf_();
    double a = 1.0, b = sqrt(1.0 - k * k);
    for (int it = 0; it < 25; ++it) {
        double an = (a + b) / 2.0;
        double bn = sqrt(a * b);
        a = an;
        b = bn;
        if (fabs(a - b) < kEps) break;
    }
    return kPi / (2.0 * a);
}
SAF_DEVFN double elliptic_e_ref(double k) {
    double ka = fabs(k);
    if (isnan(ka) || ka > 1.0) return nan_();
    double a = 1.0, b = sqrt(1.0 - k * k);
    double k2 = k * k;
    double sum = 1.0 - k2 / 2.0;
    double pow2 = 0.5;
    for (int it = 0; it < 25; ++it) {
        double an = (a + b) / 2.0;
        double bn = sqrt(a * b);
        double cn = (a - b) / 2.0;
        sum -= pow2 * cn * cn;
        a = an;
        b = bn;
        pow2 *= 2.0;
        if (fabs(cn) < kEps) break;
    }
    return kPi / (2.0 * a) * sum;
}

// ---- bessel.rs ----
template <int N>
SAF_DEV double horner(double x, const double (&c)[N]) { // bessel.rs:693-701: c[0] + x(c[1] + ...)
    double s = 0.0;
#pragma unroll
    for (int i = N - 1; i >= 0; --i) s = s * x + c[i];
    return s;
}
SAF_DEVFN double bessel_j0(double x) {
    const double num[6] = {57568490574.0, -13362590354.0, 651619640.7, -11214424.18, 77392.33017, -184.9052456};
    const double den[6] = {57568490411.0, 1029532985.0, 9494680.718, 59272.64853, 267.8532712, 1.0};
    const double pc[5] = {1.0, -0.1098628627e-2, 0.2734510407e-4, -0.2073370639e-5, 0.2093887211e-6};
    const double ps[5] = {-0.1562499995e-1, 0.1430488765e-3, -0.6911147651e-5, 0.7621095161e-6, 0.934935152e-7};
    double ax = fabs(x);
    if (ax < 8.0) {
        double y = x * x;
        return horner(y, num) / horner(y, den);
    }
    double z = 8.0 / ax, y = z * z, xx = ax - 0.785398164;
    double ts = sqrt(kFrac2Pi / ax);
    return ts * (cos(xx) * horner(y, pc) - z * sin(xx) * horner(y, ps));
}
SAF_DEVFN double bessel_j1(double x) {
    const double num[6] = {72362614232.0, -7895059235.0, 242396853.1, -2972611.439, 15704.48260, -30.16036606};
    const double den[6] = {144725228442.0, 2300535178.0, 18583304.74, 99447.43394, 376.9991397, 1.0};
    const double pc[5] = {1.0, 0.183105e-2, -0.3516396496e-4, 0.2457520174e-5, -0.240337019e-6};
    const double ps[5] = {0.04687499995, -0.2002690873e-3, 0.8449199096e-5, -0.88228987e-6, 0.105787412e-6};
    double ax = fabs(x);
    if (ax < 8.0) {
        double y = x * x;
        return x * (horner(y, num) / horner(y, den));
    }
    double z = 8.0 / ax, y = z * z, xx = ax - 2.356194491;
    double ts = sqrt(kFrac2Pi / ax);
    double ans = ts * (cos(xx) * horner(y, pc) - z * sin(xx) * horner(y, ps));
    return x < 0.0 ? -ans : ans;
}
SAF_DEV int sat_i32(double v) {
    if (isnan(v)) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int)v;
}
SAF_DEVFN double bessel_j_ref(int n, double x) { // bessel.rs:16-141
    const int na = n < 0 ? -n : n;
    const double ax = fabs(x);
    if (ax < 1e-10) return na == 0 ? 1.0 : 0.0;
    const bool flip = n < 0 && (na % 2 == 1);
    if ((double)na <= ax) { // forward recurrence
        double j0 = bessel_j0(x);
        if (na == 0) return j0;
        double j1 = bessel_j1(x);
        if (na == 1) return n < 0 ? -j1 : j1;
        double jp = j0, jc = j1, k_t = 1.0;
        for (int i = 1; i < na; ++i) {
            double jn = (2.0 * k_t / x) * jc - jp;
            jp = jc;
            jc = jn;
            k_t += 1.0;
        }
        return flip ? -jc : jc;
    }
    // Miller backward recurrence
    const int n_start = na + sat_i32(round(sqrt(50.0 * (double)na))) + 15;
    double j_next = 0.0, j_curr = 1e-30, result = 0.0, sum = 0.0, comp = 0.0;
    double k_t = (double)n_start;
    for (int k = n_start; k >= 0; --k) {
        double j_prev = (2.0 * k_t / x) * j_curr - j_next;
        if (k == na) result = j_curr;
        if (k == 0 || (k % 2 == 0)) {
            double term = (k == 0) ? j_curr : 2.0 * j_curr;
            double y = term - comp;
            double t = sum + y;
            comp = (t - sum) - y;
            sum = t;
        }
        j_next = j_curr;
        j_curr = j_prev;
        k_t -= 1.0;
    }
    double r = result * (1.0 / sum);
    return flip ? -r : r;
}
SAF_DEVFN double bessel_y0(double x) {
    const double num[6] = {-2957821389.0, 7062834065.0, -512359803.6, 10879881.29, -86327.92757, 228.4622733};
    const double den[6] = {40076544269.0, 745249964.8, 7189466.438, 47447.26470, 226.1030244, 1.0};
    const double ps[5] = {1.0, -0.1098628627e-2, 0.2734510407e-4, -0.2073370639e-5, 0.2093887211e-6};
    const double pc[5] = {-0.1562499995e-1, 0.1430488765e-3, -0.6911147651e-5, 0.7621095161e-6, 0.934935152e-7};
    if (x < 8.0) {
        double y = x * x;
        double term = horner(y, num) / horner(y, den);
        return term + kFrac2Pi * bessel_j0(x) * log(x);
    }
    double z = 8.0 / x, y = z * z, xx = x - 0.785398164;
    double ts = sqrt(kFrac2Pi / x);
    return ts * (sin(xx) * horner(y, ps) + z * cos(xx) * horner(y, pc));
}
SAF_DEVFN double bessel_y1(double x) {
    const double num[6] = {-0.4900604943e13, 0.1275274390e13, -0.5153438139e11, 0.7349264551e9, -0.4237922726e7, 0.8511937935e4};
    const double den[7] = {0.2499580570e14, 0.4244419664e12, 0.3733650367e10, 0.2245904002e8, 0.1020426050e6, 0.3549632885e3, 1.0};
    const double ps[5] = {1.0, 0.183105e-2, -0.3516396496e-4, 0.2457520174e-5, -0.240337019e-6};
    const double pc[5] = {0.04687499995, -0.2002690873e-3, 0.8449199096e-5, -0.88228987e-6, 0.105787412e-6};
    if (x < 8.0) {
        double y = x * x;
        double term = x * (horner(y, num) / horner(y, den));
        return term + kFrac2Pi * (bessel_j1(x) * log(x) - 1.0 / x);
    }
    double z = 8.0 / x, y = z * z, xx = x - 2.356194491;
    double ts = sqrt(kFrac2Pi / x);
    return ts * (sin(xx) * horner(y, ps) + z * cos(xx) * horner(y, pc));
}
SAF_DEVFN double bessel_y_ref(int n, double x) { // bessel.rs:256-281
    if (x <= 0.0) return nan_();
    const int na = n < 0 ? -n : n;
    double y0 = bessel_y0(x);
    if (na == 0) return y0;
    double y1 = bessel_y1(x);
    if (na == 1) return n < 0 ? -y1 : y1;
    double yp = y0, yc = y1, k_t = 1.0;
    for (int i = 1; i < na; ++i) {
        double yn = (2.0 * k_t / x) * yc - yp;
        yp = yc;
        yc = yn;
        k_t += 1.0;
    }
    return (n < 0 && na % 2 == 1) ? -yc : yc;
}
SAF_DEVFN double bessel_i0(double x) {
    const double sm[7] = {1.0, 3.5156229, 3.0899424, 1.2067492, 0.2659732, 0.0360768, 0.0045813};
    const double lg[9] = {0.39894228, 0.01328592, 0.00225319, -0.00157565, 0.00916281, -0.02057706, 0.02635537, -0.01647633, 0.00392377};
    double ax = fabs(x);
    if (ax < 3.75) {
        double q = x / 3.75;
        return horner(q * q, sm);
    }
    double y = 3.75 / ax;
    return (exp(ax) / sqrt(ax)) * horner(y, lg);
}
SAF_DEVFN double bessel_i1(double x) {
    const double sm[7] = {0.5, 0.87890594, 0.51498869, 0.15084934, 0.02658733, 0.00301532, 0.00032411};
    const double lg[9] = {0.39894228, -0.03988024, -0.00362018, 0.00163801, -0.01031555, 0.02282967, -0.02895312, 0.01787654, -0.00420059};
    double ax = fabs(x), ans;
    if (ax < 3.75) {
        double q = x / 3.75;
        ans = ax * horner(q * q, sm);
    } else {
        double y = 3.75 / ax;
        ans = (exp(ax) / sqrt(ax)) * horner(y, lg);
    }
    return x < 0.0 ? -ans : ans;
}
SAF_DEVFN double bessel_i_ref(int n, double x) { // bessel.rs:410-477
    const int na = n < 0 ? -n : n;
    if (na == 0) return bessel_i0(x);
    if (na == 1) return bessel_i1(x);
    if (fabs(x) < 1e-10) return 0.0;
    int n_start = na + sat_i32(sqrt((double)(40 * na))) + 10;
    if (n_start < na + 20) n_start = na + 20;
    double i_next = 0.0, i_curr = 1e-30, result = 0.0, sum = 0.0;
    double k_t = (double)n_start;
    for (int k = n_start; k >= 0; --k) {
        double i_prev = (2.0 * k_t / x) * i_curr + i_next;
        if (k == na) result = i_curr;
        if (k == 0) sum += i_curr;
        else if (k % 2 == 0) sum += 2.0 * i_curr;
        i_next = i_curr;
        i_curr = i_prev;
        k_t -= 1.0;
    }
    return result * (bessel_i0(x) / sum);
}
SAF_DEVFN double bessel_k0(double x) {
    const double c[7] = {-0.57721566, 0.42278420, 0.23069756, 0.03488590, 0.00262698, 0.00010750, 0.0000074};
    const double cl[8] = {1.25331414, -0.07832358, 0.02189568, -0.01062446, 0.00587872, -0.00251540, 0.00053208, -0.000025200};
    if (x <= 2.0) {
        double y = x * x / 4.0;
        double ln_term = -log(x / 2.0) * bessel_i0(x);
        return ln_term + horner(y, c);
    }
    double y = 2.0 / x;
    return (exp(-x) / sqrt(x)) * horner(y, cl);
}
SAF_DEVFN double bessel_k1(double x) {
    const double c[7] = {1.0, 0.15443144, -0.67278579, -0.18156897, -0.01919402, -0.00110404, -0.00004686};
    const double cl[8] = {1.25331414, 0.23498619, -0.03655620, 0.01504268, -0.00780353, 0.00325614, -0.00068245, 0.0000316};
    if (x <= 2.0) {
        double y = x * x / 4.0;
        double term1 = log(x) * bessel_i1(x);
        double term2 = 1.0 / x;
        return term1 + term2 * horner(y, c);
    }
    double y = 2.0 / x;
    return (exp(-x) / sqrt(x)) * horner(y, cl);
}
SAF_DEVFN double bessel_k_ref(int n, double x) { // bessel.rs:565-590
    if (x <= 0.0) return nan_();
    const int na = n < 0 ? -n : n;
    double k0 = bessel_k0(x);
    if (na == 0) return k0;
    double k1 = bessel_k1(x);
    if (na == 1) return k1;
    double kp = k0, kc = k1, k_t = 1.0;
    for (int i = 1; i < na; ++i) {
        double kn = kp + (2.0 * k_t / x) * kc;
        kp = kc;
        kc = kn;
        k_t += 1.0;
    }
    return kc;
}

// ---- polynomials.rs ----
SAF_DEVFN double hermite_ref(int n, double x) {
    if (n < 0) return nan_();
    if (n == 0) return 1.0;
    double term1 = 2.0 * x;
    if (n == 1) return term1;
    double h0 = 1.0, h1 = term1;
    for (int k = 1; k < n; ++k) {
        double h2 = (2.0 * x * h1) - (2.0 * (double)k * h0);
        h0 = h1;
        h1 = h2;
    }
    return h1;
}
SAF_DEVFN double assoc_legendre_ref(int l, int m, double x) {
    const int ma = m < 0 ? -m : m;
    if (l < 0 || ma > l) return nan_();
    double xa = fabs(x);
    if (isnan(xa) || xa > 1.0) return nan_();
    double pmm = 1.0;
    if (ma > 0) {
        double sqx = sqrt(1.0 - x * x);
        double fact = 1.0;
        for (int i = 1; i <= ma; ++i) {
            pmm = pmm * (-fact) * sqx;
            fact += 2.0;
        }
    }
    if (l == ma) return pmm;
    double pmmp1 = x * (double)(2 * ma + 1) * pmm;
    if (l == ma + 1) return pmmp1;
    double pll = 0.0, prev = pmm, curr = pmmp1;
    for (int ll = ma + 2; ll <= l; ++ll) {
        double denom = (double)ll - (double)ma;
        pll = (x * (double)(2 * ll - 1) * curr - (double)(ll + ma - 1) * prev) / denom;
        prev = curr;
        curr = pll;
    }
    return pll;
}
SAF_DEVFN double spherical_harmonic_ref(int l, int m, double theta, double phi) {
    const int ma = m < 0 ? -m : m;
    if (l < 0 || ma > l) return nan_();
    double ct = cos(theta);
    if (isnan(ct)) return nan_();
    double plm = assoc_legendre_ref(l, m, ct);
    double f_lm = 1.0;
    for (int i = 1; i <= l - ma; ++i) f_lm *= (double)i;
    double f_lpm = 1.0;
    for (int i = 1; i <= l + ma; ++i) f_lpm *= (double)i;
    double norm_sq = ((double)(2 * l + 1) / (4.0 * kPi)) * (f_lm / f_lpm);
    double norm = sqrt(norm_sq);
    return norm * plm * cos((double)m * phi);
}

// ---- builtins.rs:38-86 ----
SAF_DEVFN double builtin1(uint32_t fn, double x) {
    switch (fn) {
    case FN_TAN: return tan(x);
    case FN_COT: return 1.0 / tan(x);
    case FN_SEC: return 1.0 / cos(x);
    case FN_CSC: return 1.0 / sin(x);
    case FN_ASIN: return asin(x);
    case FN_ACOS: return acos(x);
    case FN_ATAN: return atan(x);
    case FN_ACOT: return kFracPi2 - atan(x);
    case FN_ASEC: return acos(1.0 / x);
    case FN_ACSC: return asin(1.0 / x);
    case FN_SINH: return sinh(x);
    case FN_COSH: return cosh(x);
    case FN_TANH: return tanh(x);
    case FN_COTH: return 1.0 / tanh(x);
    case FN_SECH: return 1.0 / cosh(x);
    case FN_CSCH: return 1.0 / sinh(x);
    case FN_ASINH: return rs_asinh(x);
    case FN_ACOSH: return rs_acosh(x);
    case FN_ATANH: return rs_atanh(x);
    case FN_ACOTH: return rs_atanh(1.0 / x);
    case FN_ACSCH: return rs_asinh(1.0 / x);
    case FN_ASECH: return rs_acosh(1.0 / x);
    case FN_EXPM1: return expm1(x);
    case FN_EXPNEG: return exp(-x);
    case FN_LOG1P: return log1p(x);
    case FN_CBRT: return cbrt(x);
    case FN_ABS: return fabs(x);
    case FN_SIGNUM: return signum(x);
    case FN_FLOOR: return floor(x);
    case FN_CEIL: return ceil(x);
    case FN_ROUND: return round(x);
    case FN_ERF: return erf_ref(x);
    case FN_ERFC: return erfc_ref(x);
    case FN_GAMMA: return gamma_ref(x);
    case FN_LGAMMA: return lgamma_ref(x);
    case FN_DIGAMMA: return digamma_ref(x);
    case FN_TRIGAMMA: return trigamma_ref(x);
    case FN_TETRAGAMMA: return tetragamma_ref(x);
    case FN_SINC: return sinc(x);
    case FN_LAMBERTW: return lambert_w_ref(x);
    case FN_ELLIPTICK: return elliptic_k_ref(x);
    case FN_ELLIPTICE: return elliptic_e_ref(x);
    case FN_ZETA: return zeta_ref(x);
    case FN_EXPPOLAR: return exp(x);
    default: return nan_();
    }
}

// ---- builtins.rs:90-115 ----
SAF_DEVFN double builtin2(uint32_t fn, double x1, double x2) {
    int n;
    switch (fn) {
    case FN_ATAN2: return atan2(x1, x2);
    case FN_LOG:
        if (x1 <= 0.0 || x1 == 1.0 || x2 < 0.0) return nan_();
        return log(x2) / log(x1);
    case FN_BESSELJ: return round_to_i32(x1, n) ? bessel_j_ref(n, x2) : nan_();
    case FN_BESSELY: return round_to_i32(x1, n) ? bessel_y_ref(n, x2) : nan_();
    case FN_BESSELI: return round_to_i32(x1, n) ? bessel_i_ref(n, x2) : nan_();
    case FN_BESSELK: return round_to_i32(x1, n) ? bessel_k_ref(n, x2) : nan_();
    case FN_POLYGAMMA: return round_to_i32(x1, n) ? polygamma_ref(n, x2) : nan_();
    case FN_BETA: return beta_ref(x1, x2);
    case FN_ZETADERIV: return round_to_i32(x1, n) ? zeta_deriv_ref(n, x2) : nan_();
    case FN_HERMITE: return round_to_i32(x1, n) ? hermite_ref(n, x2) : nan_();
    default: return nan_();
    }
}

// ---- builtins.rs:119-139 ----
SAF_DEVFN double builtin3(uint32_t fn, double x1, double x2, double x3) {
    int l, m;
    if (fn == FN_ASSOCLEGENDRE && round_to_i32(x1, l) && round_to_i32(x2, m)) return assoc_legendre_ref(l, m, x3);
    return nan_();
}
SAF_DEVFN double builtin4(uint32_t fn, double x1, double x2, double x3, double x4) {
    int l, m;
    if (fn == FN_SPHERICALHARMONIC && round_to_i32(x1, l) && round_to_i32(x2, m))
        return spherical_harmonic_ref(l, m, x3, x4);
    return nan_();
}

} // namespace dm
} // namespace saf
