/* Internal to oracle/: special functions restated from
 * /root/reference/src/math/logic/functions/ (all .rs files) with T = f64.
 * TEST INFRASTRUCTURE ONLY (see saf_oracle.h). */
#ifndef SAF_ORACLE_MATH_H
#define SAF_ORACLE_MATH_H
#include <stdint.h>

double saf_o_powi(double a, int32_t b);            /* compiler-rt __powidf2 */
double saf_o_signum(double x);                      /* Rust f64::signum */
double saf_o_asinh(double x);                       /* Rust std formulas */
double saf_o_acosh(double x);
double saf_o_atanh(double x);

double saf_o_erf(double x);                         /* erf.rs:9-51 */
double saf_o_erfc(double x);                        /* erf.rs:64-126 */
double saf_o_gamma(double x);                       /* gamma.rs:33-65 */
double saf_o_lgamma(double x);                      /* gamma.rs:84-103 */
double saf_o_beta(double x, double y);              /* beta.rs:11-33 */
double saf_o_digamma(double x);                     /* polygamma.rs:11-41 */
double saf_o_trigamma(double x);                    /* polygamma.rs:49-68 */
double saf_o_tetragamma(double x);                  /* polygamma.rs:75-94 */
double saf_o_polygamma(int32_t n, double x);        /* polygamma.rs:102-208 */
double saf_o_zeta(double x);                        /* zeta.rs:15-89 */
double saf_o_zeta_deriv(int32_t n, double x);       /* zeta.rs:186-266 */
double saf_o_lambert_w(double x);                   /* lambert_w.rs:13-91 */
double saf_o_elliptic_k(double k);                  /* elliptic.rs:11-40 */
double saf_o_elliptic_e(double k);                  /* elliptic.rs:48-78 */
double saf_o_bessel_j(int32_t n, double x);         /* bessel.rs:16-35 */
double saf_o_bessel_y(int32_t n, double x);         /* bessel.rs:256-281 */
double saf_o_bessel_i(int32_t n, double x);         /* bessel.rs:410-477 */
double saf_o_bessel_k(int32_t n, double x);         /* bessel.rs:565-590 */
double saf_o_hermite(int32_t n, double x);          /* polynomials.rs:10-31 */
double saf_o_assoc_legendre(int32_t l, int32_t m, double x); /* polynomials.rs:39-88 */
double saf_o_spherical_harmonic(int32_t l, int32_t m, double theta, double phi); /* polynomials.rs:96-128 */
#endif
