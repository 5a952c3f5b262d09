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
// reproduced.  rustc never contracts a*b+c, and only the MulAdd-family opcodes may fuse -- but the
// CUDA math library NEEDS nvcc's default -fmad=true (its sin/cos range reduction is written as a*b+c
// and loses ~9 bits when not fused).  So the translation unit keeps -fmad=true and every
// reference-mirroring computation below is written on the `E` ("exactly rounded") wrapper type, whose
// operators are the explicit-rounding intrinsics __dadd_rn/__dmul_rn/__ddiv_rn that nvcc never fuses.
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
// Work caps (documented deviation): the reference's unbounded recurrences (`while xv < 6`)
// and order-n loops would hang a GPU (trigamma(-inf) never returns on the CPU either).  NaN is
// returned when more than kMaxShift recurrence steps or an order above kMaxOrder would be needed.
constexpr double kMaxShift = 4194304.0;
constexpr int kMaxOrder = 65536;

SAF_DEV double nan_() { return __longlong_as_double(0x7ff8000000000000LL); }
SAF_DEV double inf_() { return __longlong_as_double(0x7ff0000000000000LL); }

// A double whose + - * / are individually rounded and can never be contracted into an FMA.
struct E {
    double v;
    E() = default;
    SAF_DEV E(double x) : v(x) {}
    SAF_DEV E(int x) : v((double)x) {}
};
SAF_DEV E operator+(E a, E b) { return E(__dadd_rn(a.v, b.v)); }
SAF_DEV E operator-(E a, E b) { return E(__dsub_rn(a.v, b.v)); }
SAF_DEV E operator*(E a, E b) { return E(__dmul_rn(a.v, b.v)); }
SAF_DEV E operator/(E a, E b) { return E(__ddiv_rn(a.v, b.v)); }
SAF_DEV E operator-(E a) { return E(-a.v); }
SAF_DEV E &operator+=(E &a, E b) { a = a + b; return a; }
SAF_DEV E &operator-=(E &a, E b) { a = a - b; return a; }
SAF_DEV E &operator*=(E &a, E b) { a = a * b; return a; }
SAF_DEV bool operator<(E a, E b) { return a.v < b.v; }
SAF_DEV bool operator>(E a, E b) { return a.v > b.v; }
SAF_DEV bool operator<=(E a, E b) { return a.v <= b.v; }
SAF_DEV bool operator>=(E a, E b) { return a.v >= b.v; }
SAF_DEV bool operator==(E a, E b) { return a.v == b.v; }
SAF_DEV bool operator!=(E a, E b) { return a.v != b.v; }
// CUDA math library on E
SAF_DEV E sin(E x) { return E(::sin(x.v)); }
SAF_DEV E cos(E x) { return E(::cos(x.v)); }
SAF_DEV E tan(E x) { return E(::tan(x.v)); }
SAF_DEV E exp(E x) { return E(::exp(x.v)); }
SAF_DEV E log(E x) { return E(::log(x.v)); }
SAF_DEV E log1p(E x) { return E(::log1p(x.v)); }
SAF_DEV E pow(E x, E y) { return E(::pow(x.v, y.v)); }
SAF_DEV E sqrt(E x) { return E(__dsqrt_rn(x.v)); }
SAF_DEV E fabs(E x) { return E(::fabs(x.v)); }
SAF_DEV E round(E x) { return E(::round(x.v)); }
SAF_DEV E trunc(E x) { return E(::trunc(x.v)); }
SAF_DEV E fmax(E x, E y) { return E(::fmax(x.v, y.v)); }
SAF_DEV E hypot(E x, E y) { return E(::hypot(x.v, y.v)); }
SAF_DEV E copysign(E x, E y) { return E(::copysign(x.v, y.v)); }
SAF_DEV E asin(E x) { return E(::asin(x.v)); }
SAF_DEV E acos(E x) { return E(::acos(x.v)); }
SAF_DEV E atan(E x) { return E(::atan(x.v)); }
SAF_DEV E atan2(E y, E x) { return E(::atan2(y.v, x.v)); }
SAF_DEV E sinh(E x) { return E(::sinh(x.v)); }
SAF_DEV E cosh(E x) { return E(::cosh(x.v)); }
SAF_DEV E tanh(E x) { return E(::tanh(x.v)); }
SAF_DEV E expm1(E x) { return E(::expm1(x.v)); }
SAF_DEV E cbrt(E x) { return E(::cbrt(x.v)); }
SAF_DEV E floor(E x) { return E(::floor(x.v)); }
SAF_DEV E ceil(E x) { return E(::ceil(x.v)); }
SAF_DEV bool isnan(E x) { return ::isnan(x.v); }
SAF_DEV bool isinf(E x) { return ::isinf(x.v); }
SAF_DEV bool signbit(E x) { return ::signbit(x.v); }

// f64::powi -> compiler-rt __powidf2: square-and-multiply from the low bit, reciprocal last
SAF_DEV E powi(E a, int b) {
    const bool recip = b < 0;
    E r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

SAF_DEV E signum(E x) { return isnan(x) ? x : copysign(1.0, x); }
SAF_DEV E fract(E x) { return x - trunc(x); }

// Rust std f64::asinh / acosh / atanh formulas
SAF_DEV E rs_asinh(E x) {
    E ax = fabs(x);
    E ix = 1.0 / ax;
    return copysign(log1p(ax + (ax / (hypot(1.0, ix) + ix))), x);
}
SAF_DEV E rs_acosh(E x) {
    if (x < 1.0) return nan_();
    return log(x + (sqrt(x - 1.0) * sqrt(x + 1.0)));
}
SAF_DEV E rs_atanh(E x) { return 0.5 * log1p((2.0 * x) / (1.0 - x)); }

SAF_DEV E sinc(E x) { return fabs(x) < kEps ? 1.0 : sin(x) / x; }

SAF_DEV bool round_to_i32(E x, int &out) {
    E r = round(x);
    if (!(r >= -2147483648.0 && r <= 2147483647.0)) return false;
    out = (int)r.v;
    return true;
}

// ---- math/logic/functions/helpers.rs ----
SAF_DEV E sign_from_delta(E d) { return isnan(d) ? nan_() : (signbit(d) ? -1.0 : 1.0); }
SAF_DEV E signed_infinity(E s) { return isnan(s) ? nan_() : (signbit(s) ? -inf_() : inf_()); }
SAF_DEV long long to_i64_or0(E v) {
    if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return 0;
    return (long long)v.v;
}
SAF_DEV E gamma_pole_sign(E x) { return (to_i64_or0(-x) & 1) == 0 ? 1.0 : -1.0; }
SAF_DEV E gamma_sign(E x) {
    if (x > 0.0) return 1.0;
    return (to_i64_or0(-x) & 1) == 0 ? -1.0 : 1.0;
}

// ---- erf.rs:9-51: alternating Taylor series, Kahan-compensated, at most 30 terms ----
SAF_DEVFN E erf_ref(E x0) {
    const E sign = signum(x0);
    const E x = fabs(x0);
    const E coeff = 2.0 / sqrt(kPi);
    const E xx = x * x;
    E sum = 0.0, comp = 0.0, fact = 1.0, power = x;
    for (int n = 0; n < 30; ++n) {
        E term = power / (fact * (E)(2 * n + 1));
        if (isnan(term) || isinf(term)) break;
        E st = (n & 1) ? -term : term;
        E y = st - comp;
        E t = sum + y;
        comp = (t - sum) - y;
        sum = t;
        fact *= (E)(n + 1);
        power *= xx;
        if (fabs(term) < DBL_EPSILON) break;
    }
    return sign * coeff * sum;
}

// ---- erf.rs:64-126 ----
SAF_DEVFN E erfc_ref(E x) {
    if (isnan(x)) return x;
    if (isinf(x)) return signbit(x) ? 2.0 : 0.0;
    const E ax = fabs(x);
    if (x > 28.0) return 0.0;
    if (ax < 1.5) return 1.0 - erf_ref(x);
    // continued fraction (modified Lentz) on |x| >= 1.5; the reference recurses once for x < 0
    const E tiny = 1e-30;
    E f = ax;
    if (fabs(f) < tiny) f = tiny;
    E c = f, d = 0.0;
    for (int j = 1; j < 200; ++j) {
        E a = 0.5 * (E)j;
        d = ax + a * d;
        if (fabs(d) < tiny) d = tiny;
        d = 1.0 / d;
        c = ax + a / c;
        if (fabs(c) < tiny) c = tiny;
        E delta = c * d;
        f *= delta;
        if (fabs(delta - 1.0) < DBL_EPSILON) break;
    }
    E pos = (exp(-ax * ax) / sqrt(kPi)) / f;
    return x < 0.0 ? 2.0 - pos : pos;
}

// ---- gamma.rs ----
SAF_DEV E lanczos_ag(E xm1) {
    const E c[9] = {0.9999999999998099,   676.5203681218851,    -1259.1392167224028,
                         771.3234287776531,    -176.6150291621406,   12.507343278686905,
                         -0.13857109526572012, 9.984369578019572e-6, 1.5056327351493116e-7};
    E ag = c[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) ag += c[i] / (xm1 + (E)i);
    return ag;
}
SAF_DEV E gamma_lanczos(E x) { // x >= 0.5, gamma.rs:57-64
    E xm1 = x - 1.0;
    E ag = lanczos_ag(xm1);
    E t = xm1 + 7.0 + 0.5;
    return sqrt(2.0 * kPi) * pow(t, xm1 + 0.5) * exp(-t) * ag;
}
SAF_DEVFN E gamma_ref(E x) { // gamma.rs:33-65
    if (x <= 0.0 && fract(x) == 0.0) {
        E delta = x - round(x);
        return signed_infinity(gamma_pole_sign(x) * sign_from_delta(delta));
    }
    if (x < 0.5) {
        // reflection; 1 - x > 0.5 so the recursion of the reference has depth one -- except that
        // 1 - x can round to a non-positive integer pole check, which the inner call repeats
        E y = 1.0 - x;
        E gy;
        if (y <= 0.0 && fract(y) == 0.0) {
            E delta = y - round(y);
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
SAF_DEV E lgamma_pos(E x) { // x >= 0.5, gamma.rs:94-102
    E xm1 = x - 1.0;
    E ag = lanczos_ag(xm1);
    E t = xm1 + 7.0 + 0.5;
    return log(sqrt(2.0 * kPi)) + (xm1 + 0.5) * log(t) - t + log(fabs(ag));
}
SAF_DEVFN E lgamma_ref(E x) { // gamma.rs:84-103
    if (x <= 0.0 && fract(x) == 0.0) return inf_();
    if (x < 0.5) {
        E sin_term = fabs(sin(kPi * x));
        E y = 1.0 - x;
        E ly = (y <= 0.0 && fract(y) == 0.0) ? inf_() : lgamma_pos(y);
        return log(kPi) - log(sin_term) - ly;
    }
    return lgamma_pos(x);
}

// ---- beta.rs:11-33 ----
SAF_DEVFN E beta_ref(E x, E y) {
    if (x <= 0.0 && fract(x) == 0.0) return signed_infinity(gamma_sign(x));
    if (y <= 0.0 && fract(y) == 0.0) return signed_infinity(gamma_sign(y));
    E xy = x + y;
    if (xy <= 0.0 && fract(xy) == 0.0) return 0.0;
    E sign = gamma_sign(x) * gamma_sign(y) * gamma_sign(xy);
    E lbeta = lgamma_ref(x) + lgamma_ref(y) - lgamma_ref(xy);
    return sign * exp(lbeta);
}

// ---- polygamma.rs ----
SAF_DEV E digamma_pos(E x) { // x >= 0.5, polygamma.rs:25-40
    E xv = x, result = 0.0;
    while (xv < 6.0) {
        result -= 1.0 / xv;
        xv += 1.0;
    }
    result += log(xv) - 0.5 / xv;
    E x2 = xv * xv;
    E t1 = 1.0 / (12.0 * x2);
    E t2 = 1.0 / (120.0 * x2 * x2);
    E t3 = 1.0 / (252.0 * x2 * x2 * x2);
    return result - t1 + t2 - t3;
}
SAF_DEVFN E digamma_ref(E x) { // polygamma.rs:11-41
    if (x <= 0.0 && fract(x) == 0.0) {
        E delta = x - round(x);
        return signed_infinity(-sign_from_delta(delta));
    }
    if (x < 0.5) {
        E y = 1.0 - x;
        E dy;
        if (y <= 0.0 && fract(y) == 0.0) {
            E delta = y - round(y);
            dy = signed_infinity(-sign_from_delta(delta));
        } else {
            dy = digamma_pos(y);
        }
        return dy - kPi * cos(kPi * x) / sin(kPi * x);
    }
    return digamma_pos(x);
}
SAF_DEVFN E trigamma_ref(E x) { // polygamma.rs:49-68
    if (x <= 0.0 && fract(x) == 0.0) return inf_();
    if (x < -kMaxShift) return nan_();
    E xv = x, r = 0.0;
    while (xv < 6.0) {
        r += 1.0 / (xv * xv);
        xv += 1.0;
    }
    E x2 = xv * xv;
    return r + 1.0 / xv + 0.5 / x2 + 1.0 / (6.0 * x2 * xv) - 1.0 / (30.0 * x2 * x2 * xv) +
           1.0 / (42.0 * x2 * x2 * x2 * xv);
}
SAF_DEVFN E tetragamma_ref(E x) { // polygamma.rs:75-94
    if (x <= 0.0 && fract(x) == 0.0) {
        E delta = x - round(x);
        return signed_infinity(-sign_from_delta(delta));
    }
    if (x < -kMaxShift) return nan_();
    E xv = x, r = 0.0;
    while (xv < 6.0) {
        r -= 2.0 / (xv * xv * xv);
        xv += 1.0;
    }
    E x2 = xv * xv;
    return r - 1.0 / x2 + 1.0 / (x2 * xv) + 1.0 / (2.0 * x2 * x2) + 1.0 / (6.0 * x2 * x2 * xv);
}
SAF_DEVFN E polygamma_ref(int n, E x) { // polygamma.rs:102-208
    if (n < 0) return nan_();
    if (n == 0) return digamma_ref(x);
    if (n == 1) return trigamma_ref(x);
    if (n > kMaxOrder || x < -kMaxShift) return nan_();
    if (x <= 0.0 && fract(x) == 0.0) {
        E delta = x - round(x);
        E sign = (n % 2 == 0) ? -sign_from_delta(delta) : 1.0;
        return signed_infinity(sign);
    }
    E xv = x, r = 0.0;
    const E sign = ((n + 1) % 2 == 0) ? 1.0 : -1.0;
    E factorial = 1.0;
    for (int i = 1; i <= n; ++i) factorial *= (E)i;
    const int np1 = n + 1;
    while (xv < 15.0) {
        r += sign * factorial / powi(xv, np1);
        xv += 1.0;
    }
    const E asym_sign = (n % 2 == 0) ? -1.0 : 1.0;
    const E bnum[5] = {1.0, -1.0, 1.0, -1.0, 5.0};
    const E bden[5] = {6.0, 30.0, 42.0, 30.0, 66.0};
    E nm1_fact = 1.0;
    for (int i = 1; i < n; ++i) nm1_fact *= (E)i;
    E sum = nm1_fact / powi(xv, n);
    sum += factorial / (2.0 * powi(xv, np1));
    E xpow = powi(xv, n + 2);
    E fact_ratio = factorial * (E)(n + 1);
    E prev = DBL_MAX;
    for (int k = 0; k < 5; ++k) {
        int two_k = 2 * (k + 1);
        E f2k = 1.0;
        for (int i = 1; i <= two_k; ++i) f2k *= (E)i;
        E term = (bnum[k] / bden[k]) * fact_ratio / (f2k * xpow);
        if (fabs(term) > prev) break;
        prev = fabs(term);
        sum += term;
        xpow *= xv * xv;
        fact_ratio *= (E)(n + two_k) * (E)(n + two_k + 1);
    }
    return r + asym_sign * sum;
}

// ---- zeta.rs ----
SAF_DEVFN E zeta_borwein(E s) { // zeta.rs:107-163
    const int N = 14;
    E denom = 1.0 - pow(2.0, 1.0 - s);
    if (fabs(denom) < 1e-15) return signed_infinity(sign_from_delta(s - 1.0));
    E d[N + 1];
    const E n_t = (E)N;
    E term = 1.0 / n_t;
    E inner = term;
    d[0] = n_t * inner;
    for (int k = 1; k <= N; ++k) {
        E i = (E)(k - 1);
        term = term * 4.0 * (n_t + i) * (n_t - i) / ((E)(2 * k - 1) * (E)(2 * k));
        inner += term;
        d[k] = n_t * inner;
    }
    const E d_n = d[N];
    E sum = 0.0, comp = 0.0;
    for (int k = 0; k < N; ++k) {
        E sign = (k % 2 == 0) ? 1.0 : -1.0;
        E cur = sign * (d[k] - d_n) / pow((E)(k + 1), s);
        E y = cur - comp;
        E t = sum + y;
        comp = (t - sum) - y;
        sum = t;
    }
    return -sum / (d_n * denom);
}
SAF_DEVFN E zeta_nonneg(E x) { // zeta.rs:38-88, x >= 0 and away from the pole
    if (x > 1.0 && x <= 1.5) {
        E sum = 0.0, comp = 0.0;
        for (int k = 1; k <= 100; ++k) {
            E term = 1.0 / pow((E)k, x);
            E y = term - comp;
            E t = sum + y;
            comp = (t - sum) - y;
            sum = t;
        }
        const E n = 100.0;
        E n_pow_x = pow(n, x);
        E n_pow_1mx = pow(n, 1.0 - x);
        E em_integral = n_pow_1mx / (x - 1.0);
        E em_boundary = 0.5 / n_pow_x;
        E b2 = x / (12.0 * pow(n, x + 1.0));
        E b4 = -x * (x + 1.0) * (x + 2.0) / (720.0 * pow(n, x + 3.0));
        return sum + em_integral + em_boundary + b2 + b4;
    }
    return zeta_borwein(x);
}
SAF_DEVFN E zeta_ref(E x) { // zeta.rs:15-89
    E delta = x - 1.0;
    if (fabs(delta) < 1e-10) return signed_infinity(sign_from_delta(delta));
    if (x < 0.0) {
        E y = 1.0 - x; // > 1: the reference's recursion has depth one
        E gs = gamma_ref(y);
        E dy = y - 1.0;
        E z = (fabs(dy) < 1e-10) ? signed_infinity(sign_from_delta(dy)) : zeta_nonneg(y);
        E term1 = pow(2.0, x);
        E term2 = pow(kPi, x - 1.0);
        E term3 = sin(kPi * x / 2.0);
        return term1 * term2 * term3 * gs * z;
    }
    return zeta_nonneg(x);
}
SAF_DEVFN E zeta_reflection_base(E s) { // zeta.rs:308-326
    E oms = 1.0 - s;
    E a = pow(2.0, s);
    E b = pow(kPi, s - 1.0);
    E c = sin(kPi * s * 0.5);
    E d = gamma_ref(oms);
    E z = zeta_ref(oms);
    return a * b * c * d * z;
}
// zeta.rs:331-374.  The reference recurses for n >= 5 (two calls per level); the device walks the
// same binary recursion tree with an explicit stack of (order, point) frames: the value of a node of
// order n at point s is [node(n-1, s+h) - node(n-1, s-h)] / (2h), leaves have order <= 4.
SAF_DEV E cfd_leaf(int n, E s, E h) {
    if (n == 1) return (zeta_reflection_base(s + h) - zeta_reflection_base(s - h)) / (2.0 * h);
    if (n == 2) {
        E zp = zeta_reflection_base(s + h), zc = zeta_reflection_base(s), zm = zeta_reflection_base(s - h);
        return (zp - 2.0 * zc + zm) / (h * h);
    }
    if (n == 3) {
        E zp2 = zeta_reflection_base(s + 2.0 * h), zp1 = zeta_reflection_base(s + h);
        E zm1 = zeta_reflection_base(s - h), zm2 = zeta_reflection_base(s - 2.0 * h);
        return (-zp2 + 2.0 * zp1 - 2.0 * zm1 + zm2) / (2.0 * h * h * h);
    }
    E two_h = 2.0 * h;
    E zp2 = zeta_reflection_base(s + two_h), zp1 = zeta_reflection_base(s + h);
    E zc = zeta_reflection_base(s), zm1 = zeta_reflection_base(s - h), zm2 = zeta_reflection_base(s - two_h);
    return (zp2 - 4.0 * zp1 + 6.0 * zc - 4.0 * zm1 + zm2) / (h * h * h * h);
}
constexpr int kMaxZetaDerivDepth = 12; // orders up to 16 (zeta_deriv_ref refuses more)
SAF_DEVFN E centered_finite_diff(int n, E s, E h) {
    if (n <= 4) return cfd_leaf(n, s, h);
    int depth = n - 4;
    if (depth > kMaxZetaDerivDepth) return nan_(); // documented deviation: refuse exponential work
    // post-order evaluation of a complete binary tree of `depth` levels
    E val[kMaxZetaDerivDepth + 1];
    unsigned path = 0; // bit i = 1 when level i took the "minus" branch
    int level = 0;
    E cur_s[kMaxZetaDerivDepth + 1];
    cur_s[0] = s;
    unsigned state[kMaxZetaDerivDepth + 1]; // 0 = need plus child, 1 = need minus child, 2 = done
    for (int i = 0; i <= kMaxZetaDerivDepth; ++i) state[i] = 0;
    E plus_val[kMaxZetaDerivDepth + 1];
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
SAF_DEVFN E zeta_deriv_ref(int n, E x) { // zeta.rs:186-302
    if (n < 0) return nan_();
    if (n == 0) return zeta_ref(x);
    if (n > 16) return nan_(); // work cap: 2^(n-4) zeta evaluations below s = 1
    E delta = x - 1.0;
    if (fabs(delta) < 1e-10) {
        E sign = (n % 2 == 0) ? sign_from_delta(delta) : -1.0;
        return signed_infinity(sign);
    }
    if (x > 1.0) {
        E sum = 0.0, comp = 0.0;
        for (int k = 1; k <= 200; ++k) {
            E k_t = (E)k;
            E ln_k = log(k_t);
            E lp;
            if (n <= 5) {
                lp = 1.0;
                for (int j = 0; j < n; ++j) lp *= ln_k;
            } else {
                lp = pow(ln_k, (E)n);
            }
            E term = lp / pow(k_t, x);
            E y = term - comp;
            E t = sum + y;
            comp = (t - sum) - y;
            sum = t;
            if (k > 50 && fabs(term) < 1e-10 * 0.01) break;
        }
        return ((n % 2 == 0) ? 1.0 : -1.0) * sum;
    }
    const E h = 1e-7;
    if (n == 1) return (zeta_reflection_base(x + h) - zeta_reflection_base(x - h)) / (2.0 * h);
    if (n == 2) {
        E zp = zeta_reflection_base(x + h), zc = zeta_reflection_base(x), zm = zeta_reflection_base(x - h);
        return (zp - 2.0 * zc + zm) / (h * h);
    }
    E d_h = centered_finite_diff(n, x, h);
    E d_h2 = centered_finite_diff(n, x, h / 2.0);
    E extrapolation = powi(4.0, n) - 1.0;
    return (powi(4.0, n) * d_h2 - d_h) / extrapolation;
}

// ---- lambert_w.rs:13-91 ----
SAF_DEVFN E lambert_w_ref(E x) {
    const E e_inv = 1.0 / kE;
    if (x < -e_inv) return nan_();
    if (x == 0.0) return 0.0;
    if (fabs(x + e_inv) < 1e-12) return -1.0;
    E w;
    if (x < -0.3) {
        E arg = fmax(2.0 * (kE * x + 1.0), 0.0);
        E p = sqrt(arg);
        const E c1 = 11.0 / 72.0;
        w = -1.0 + p - p * p / 3.0 + c1 * p * p * p;
    } else if (x < 0.0) {
        E p = sqrt(2.0 * (kE * x + 1.0));
        w = -1.0 + p;
    } else if (x < 1.0) {
        w = x * (1.0 - x * (1.0 - x * 1.5));
    } else if (x < 3.0) {
        E l = log(x);
        E l_ln = log(l);
        w = l - ((l_ln > 0.0) ? l_ln : E(0.0));
    } else {
        E l1 = log(x);
        E l2 = log(l1);
        w = l1 - l2 + l2 / l1;
    }
    const E tol = 1e-15;
    for (int it = 0; it < 50; ++it) {
        if (w <= -1.0) w = -0.99;
        E ew = exp(w);
        E f = w * ew - x;
        E w1 = w + 1.0;
        if (fabs(w1) < tol) break;
        E fp = ew * w1;
        E fpp = ew * (w + 2.0);
        E d = f * fp / (fp * fp - 0.5 * f * fpp);
        w -= d;
        if (fabs(d) < tol * (1.0 + fabs(w))) break;
    }
    return w;
}

// ---- elliptic.rs ----
SAF_DEVFN E elliptic_k_ref(E k) {
    E ka = fabs(k);
    if (isnan(ka) || ka > 1.0) return nan_();
    if (ka == 1.0) return inf_();
    E a = 1.0, b = sqrt(1.0 - k * k);
    for (int it = 0; it < 25; ++it) {
        E an = (a + b) / 2.0;
        E bn = sqrt(a * b);
        a = an;
        b = bn;
        if (fabs(a - b) < kEps) break;
    }
    return kPi / (2.0 * a);
}
SAF_DEVFN E elliptic_e_ref(E k) {
    E ka = fabs(k);
    if (isnan(ka) || ka > 1.0) return nan_();
    E a = 1.0, b = sqrt(1.0 - k * k);
    E k2 = k * k;
    E sum = 1.0 - k2 / 2.0;
    E pow2 = 0.5;
    for (int it = 0; it < 25; ++it) {
        E an = (a + b) / 2.0;
        E bn = sqrt(a * b);
        E cn = (a - b) / 2.0;
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
SAF_DEV E horner(E x, const E (&c)[N]) { // bessel.rs:693-701: c[0] + x(c[1] + ...)
    E s = 0.0;
#pragma unroll
    for (int i = N - 1; i >= 0; --i) s = s * x + c[i];
    return s;
}
SAF_DEVFN E bessel_j0(E x) {
    const E num[6] = {57568490574.0, -13362590354.0, 651619640.7, -11214424.18, 77392.33017, -184.9052456};
    const E den[6] = {57568490411.0, 1029532985.0, 9494680.718, 59272.64853, 267.8532712, 1.0};
    const E pc[5] = {1.0, -0.1098628627e-2, 0.2734510407e-4, -0.2073370639e-5, 0.2093887211e-6};
    const E ps[5] = {-0.1562499995e-1, 0.1430488765e-3, -0.6911147651e-5, 0.7621095161e-6, 0.934935152e-7};
    E ax = fabs(x);
    if (ax < 8.0) {
        E y = x * x;
        return horner(y, num) / horner(y, den);
    }
    E z = 8.0 / ax, y = z * z, xx = ax - 0.785398164;
    E ts = sqrt(kFrac2Pi / ax);
    return ts * (cos(xx) * horner(y, pc) - z * sin(xx) * horner(y, ps));
}
SAF_DEVFN E bessel_j1(E x) {
    const E num[6] = {72362614232.0, -7895059235.0, 242396853.1, -2972611.439, 15704.48260, -30.16036606};
    const E den[6] = {144725228442.0, 2300535178.0, 18583304.74, 99447.43394, 376.9991397, 1.0};
    const E pc[5] = {1.0, 0.183105e-2, -0.3516396496e-4, 0.2457520174e-5, -0.240337019e-6};
    const E ps[5] = {0.04687499995, -0.2002690873e-3, 0.8449199096e-5, -0.88228987e-6, 0.105787412e-6};
    E ax = fabs(x);
    if (ax < 8.0) {
        E y = x * x;
        return x * (horner(y, num) / horner(y, den));
    }
    E z = 8.0 / ax, y = z * z, xx = ax - 2.356194491;
    E ts = sqrt(kFrac2Pi / ax);
    E ans = ts * (cos(xx) * horner(y, pc) - z * sin(xx) * horner(y, ps));
    return x < 0.0 ? -ans : ans;
}
SAF_DEV int sat_i32(E v) {
    if (isnan(v)) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int)v.v;
}
SAF_DEVFN E bessel_j_ref(int n, E x) { // bessel.rs:16-141
    if (n > kMaxOrder || n < -kMaxOrder) return nan_();
    const int na = n < 0 ? -n : n;
    const E ax = fabs(x);
    if (ax < 1e-10) return na == 0 ? 1.0 : 0.0;
    const bool flip = n < 0 && (na % 2 == 1);
    if ((E)na <= ax) { // forward recurrence
        E j0 = bessel_j0(x);
        if (na == 0) return j0;
        E j1 = bessel_j1(x);
        if (na == 1) return n < 0 ? -j1 : j1;
        E jp = j0, jc = j1, k_t = 1.0;
        for (int i = 1; i < na; ++i) {
            E jn = (2.0 * k_t / x) * jc - jp;
            jp = jc;
            jc = jn;
            k_t += 1.0;
        }
        return flip ? -jc : jc;
    }
    // Miller backward recurrence
    const int n_start = na + sat_i32(round(sqrt(50.0 * (E)na))) + 15;
    E j_next = 0.0, j_curr = 1e-30, result = 0.0, sum = 0.0, comp = 0.0;
    E k_t = (E)n_start;
    for (int k = n_start; k >= 0; --k) {
        E j_prev = (2.0 * k_t / x) * j_curr - j_next;
        if (k == na) result = j_curr;
        if (k == 0 || (k % 2 == 0)) {
            E term = (k == 0) ? j_curr : E(2.0) * j_curr;
            E y = term - comp;
            E t = sum + y;
            comp = (t - sum) - y;
            sum = t;
        }
        j_next = j_curr;
        j_curr = j_prev;
        k_t -= 1.0;
    }
    E r = result * (1.0 / sum);
    return flip ? -r : r;
}
SAF_DEVFN E bessel_y0(E x) {
    const E num[6] = {-2957821389.0, 7062834065.0, -512359803.6, 10879881.29, -86327.92757, 228.4622733};
    const E den[6] = {40076544269.0, 745249964.8, 7189466.438, 47447.26470, 226.1030244, 1.0};
    const E ps[5] = {1.0, -0.1098628627e-2, 0.2734510407e-4, -0.2073370639e-5, 0.2093887211e-6};
    const E pc[5] = {-0.1562499995e-1, 0.1430488765e-3, -0.6911147651e-5, 0.7621095161e-6, 0.934935152e-7};
    if (x < 8.0) {
        E y = x * x;
        E term = horner(y, num) / horner(y, den);
        return term + kFrac2Pi * bessel_j0(x) * log(x);
    }
    E z = 8.0 / x, y = z * z, xx = x - 0.785398164;
    E ts = sqrt(kFrac2Pi / x);
    return ts * (sin(xx) * horner(y, ps) + z * cos(xx) * horner(y, pc));
}
SAF_DEVFN E bessel_y1(E x) {
    const E num[6] = {-0.4900604943e13, 0.1275274390e13, -0.5153438139e11, 0.7349264551e9, -0.4237922726e7, 0.8511937935e4};
    const E den[7] = {0.2499580570e14, 0.4244419664e12, 0.3733650367e10, 0.2245904002e8, 0.1020426050e6, 0.3549632885e3, 1.0};
    const E ps[5] = {1.0, 0.183105e-2, -0.3516396496e-4, 0.2457520174e-5, -0.240337019e-6};
    const E pc[5] = {0.04687499995, -0.2002690873e-3, 0.8449199096e-5, -0.88228987e-6, 0.105787412e-6};
    if (x < 8.0) {
        E y = x * x;
        E term = x * (horner(y, num) / horner(y, den));
        return term + kFrac2Pi * (bessel_j1(x) * log(x) - 1.0 / x);
    }
    E z = 8.0 / x, y = z * z, xx = x - 2.356194491;
    E ts = sqrt(kFrac2Pi / x);
    return ts * (sin(xx) * horner(y, ps) + z * cos(xx) * horner(y, pc));
}
SAF_DEVFN E bessel_y_ref(int n, E x) { // bessel.rs:256-281
    if (n > kMaxOrder || n < -kMaxOrder) return nan_();
    if (x <= 0.0) return nan_();
    const int na = n < 0 ? -n : n;
    E y0 = bessel_y0(x);
    if (na == 0) return y0;
    E y1 = bessel_y1(x);
    if (na == 1) return n < 0 ? -y1 : y1;
    E yp = y0, yc = y1, k_t = 1.0;
    for (int i = 1; i < na; ++i) {
        E yn = (2.0 * k_t / x) * yc - yp;
        yp = yc;
        yc = yn;
        k_t += 1.0;
    }
    return (n < 0 && na % 2 == 1) ? -yc : yc;
}
SAF_DEVFN E bessel_i0(E x) {
    const E sm[7] = {1.0, 3.5156229, 3.0899424, 1.2067492, 0.2659732, 0.0360768, 0.0045813};
    const E lg[9] = {0.39894228, 0.01328592, 0.00225319, -0.00157565, 0.00916281, -0.02057706, 0.02635537, -0.01647633, 0.00392377};
    E ax = fabs(x);
    if (ax < 3.75) {
        E q = x / 3.75;
        return horner(q * q, sm);
    }
    E y = 3.75 / ax;
    return (exp(ax) / sqrt(ax)) * horner(y, lg);
}
SAF_DEVFN E bessel_i1(E x) {
    const E sm[7] = {0.5, 0.87890594, 0.51498869, 0.15084934, 0.02658733, 0.00301532, 0.00032411};
    const E lg[9] = {0.39894228, -0.03988024, -0.00362018, 0.00163801, -0.01031555, 0.02282967, -0.02895312, 0.01787654, -0.00420059};
    E ax = fabs(x), ans;
    if (ax < 3.75) {
        E q = x / 3.75;
        ans = ax * horner(q * q, sm);
    } else {
        E y = 3.75 / ax;
        ans = (exp(ax) / sqrt(ax)) * horner(y, lg);
    }
    return x < 0.0 ? -ans : ans;
}
SAF_DEVFN E bessel_i_ref(int n, E x) { // bessel.rs:410-477
    if (n > kMaxOrder || n < -kMaxOrder) return nan_();
    const int na = n < 0 ? -n : n;
    if (na == 0) return bessel_i0(x);
    if (na == 1) return bessel_i1(x);
    if (fabs(x) < 1e-10) return 0.0;
    int n_start = na + sat_i32(sqrt((E)(40 * na))) + 10;
    if (n_start < na + 20) n_start = na + 20;
    E i_next = 0.0, i_curr = 1e-30, result = 0.0, sum = 0.0;
    E k_t = (E)n_start;
    for (int k = n_start; k >= 0; --k) {
        E i_prev = (2.0 * k_t / x) * i_curr + i_next;
        if (k == na) result = i_curr;
        if (k == 0) sum += i_curr;
        else if (k % 2 == 0) sum += 2.0 * i_curr;
        i_next = i_curr;
        i_curr = i_prev;
        k_t -= 1.0;
    }
    return result * (bessel_i0(x) / sum);
}
SAF_DEVFN E bessel_k0(E x) {
    const E c[7] = {-0.57721566, 0.42278420, 0.23069756, 0.03488590, 0.00262698, 0.00010750, 0.0000074};
    const E cl[8] = {1.25331414, -0.07832358, 0.02189568, -0.01062446, 0.00587872, -0.00251540, 0.00053208, -0.000025200};
    if (x <= 2.0) {
        E y = x * x / 4.0;
        E ln_term = -log(x / 2.0) * bessel_i0(x);
        return ln_term + horner(y, c);
    }
    E y = 2.0 / x;
    return (exp(-x) / sqrt(x)) * horner(y, cl);
}
SAF_DEVFN E bessel_k1(E x) {
    const E c[7] = {1.0, 0.15443144, -0.67278579, -0.18156897, -0.01919402, -0.00110404, -0.00004686};
    const E cl[8] = {1.25331414, 0.23498619, -0.03655620, 0.01504268, -0.00780353, 0.00325614, -0.00068245, 0.0000316};
    if (x <= 2.0) {
        E y = x * x / 4.0;
        E term1 = log(x) * bessel_i1(x);
        E term2 = 1.0 / x;
        return term1 + term2 * horner(y, c);
    }
    E y = 2.0 / x;
    return (exp(-x) / sqrt(x)) * horner(y, cl);
}
SAF_DEVFN E bessel_k_ref(int n, E x) { // bessel.rs:565-590
    if (n > kMaxOrder || n < -kMaxOrder) return nan_();
    if (x <= 0.0) return nan_();
    const int na = n < 0 ? -n : n;
    E k0 = bessel_k0(x);
    if (na == 0) return k0;
    E k1 = bessel_k1(x);
    if (na == 1) return k1;
    E kp = k0, kc = k1, k_t = 1.0;
    for (int i = 1; i < na; ++i) {
        E kn = kp + (2.0 * k_t / x) * kc;
        kp = kc;
        kc = kn;
        k_t += 1.0;
    }
    return kc;
}

// ---- polynomials.rs ----
SAF_DEVFN E hermite_ref(int n, E x) {
    if (n > kMaxOrder) return nan_();
    if (n < 0) return nan_();
    if (n == 0) return 1.0;
    E term1 = 2.0 * x;
    if (n == 1) return term1;
    E h0 = 1.0, h1 = term1;
    for (int k = 1; k < n; ++k) {
        E h2 = (2.0 * x * h1) - (2.0 * (E)k * h0);
        h0 = h1;
        h1 = h2;
    }
    return h1;
}
SAF_DEVFN E assoc_legendre_ref(int l, int m, E x) {
    if (l > kMaxOrder) return nan_();
    const int ma = m < 0 ? -m : m;
    if (l < 0 || ma > l) return nan_();
    E xa = fabs(x);
    if (isnan(xa) || xa > 1.0) return nan_();
    E pmm = 1.0;
    if (ma > 0) {
        E sqx = sqrt(1.0 - x * x);
        E fact = 1.0;
        for (int i = 1; i <= ma; ++i) {
            pmm = pmm * (-fact) * sqx;
            fact += 2.0;
        }
    }
    if (l == ma) return pmm;
    E pmmp1 = x * (E)(2 * ma + 1) * pmm;
    if (l == ma + 1) return pmmp1;
    E pll = 0.0, prev = pmm, curr = pmmp1;
    for (int ll = ma + 2; ll <= l; ++ll) {
        E denom = (E)ll - (E)ma;
        pll = (x * (E)(2 * ll - 1) * curr - (E)(ll + ma - 1) * prev) / denom;
        prev = curr;
        curr = pll;
    }
    return pll;
}
SAF_DEVFN E spherical_harmonic_ref(int l, int m, E theta, E phi) {
    if (l > kMaxOrder) return nan_();
    const int ma = m < 0 ? -m : m;
    if (l < 0 || ma > l) return nan_();
    E ct = cos(theta);
    if (isnan(ct)) return nan_();
    E plm = assoc_legendre_ref(l, m, ct);
    E f_lm = 1.0;
    for (int i = 1; i <= l - ma; ++i) f_lm *= (E)i;
    E f_lpm = 1.0;
    for (int i = 1; i <= l + ma; ++i) f_lpm *= (E)i;
    E norm_sq = ((E)(2 * l + 1) / (4.0 * kPi)) * (f_lm / f_lpm);
    E norm = sqrt(norm_sq);
    return norm * plm * cos((E)m * phi);
}

// ---- builtins.rs:38-86 ----
SAF_DEVFN E builtin1(uint32_t fn, E x) {
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
SAF_DEVFN E builtin2(uint32_t fn, E x1, E x2) {
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
SAF_DEVFN E builtin3(uint32_t fn, E x1, E x2, E x3) {
    int l, m;
    if (fn == FN_ASSOCLEGENDRE && round_to_i32(x1, l) && round_to_i32(x2, m)) return assoc_legendre_ref(l, m, x3);
    return nan_();
}
SAF_DEVFN E builtin4(uint32_t fn, E x1, E x2, E x3, E x4) {
    int l, m;
    if (fn == FN_SPHERICALHARMONIC && round_to_i32(x1, l) && round_to_i32(x2, m))
        return spherical_harmonic_ref(l, m, x3, x4);
    return nan_();
}

} // namespace dm
} // namespace saf
