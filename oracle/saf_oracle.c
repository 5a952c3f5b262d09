/*
 * saf_oracle.c -- CPU restatement of the SymbAnaFis bytecode interpreter and
 * its batch drivers.  TEST INFRASTRUCTURE ONLY (see saf_oracle.h for the
 * reference file:line map and the parity status).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile).  glibc
 * libm supplies sin/cos/exp/log/pow/... exactly as it does for Rust's f64
 * methods on Linux.
 */
#define _GNU_SOURCE
#include "saf_oracle.h"
#include "saf_oracle_math.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define SAF_CHUNK_SIZE 256 /* batch.rs:9 */
#define SAF_FRAC_PI_2 1.57079632679489661923132169163975144
#define SAF_EPSILON 1e-14 /* src/lib.rs:321 */

/* FnOp numbering = declaration order of functions.rs:38-117 */
enum {
    FN_SIN = 0, FN_COS, FN_TAN, FN_COT, FN_SEC, FN_CSC, FN_ASIN, FN_ACOS, FN_ATAN, FN_ACOT,
    FN_ASEC, FN_ACSC, FN_SINH, FN_COSH, FN_TANH, FN_COTH, FN_SECH, FN_CSCH, FN_ASINH,
    FN_ACOSH, FN_ATANH, FN_ACOTH, FN_ACSCH, FN_ASECH, FN_EXP, FN_EXPM1, FN_EXPNEG, FN_LN,
    FN_LOG1P, FN_SQRT, FN_CBRT, FN_ABS, FN_SIGNUM, FN_FLOOR, FN_CEIL, FN_ROUND, FN_ERF,
    FN_ERFC, FN_GAMMA, FN_LGAMMA, FN_DIGAMMA, FN_TRIGAMMA, FN_TETRAGAMMA, FN_SINC,
    FN_LAMBERTW, FN_ELLIPTICK, FN_ELLIPTICE, FN_ZETA, FN_EXPPOLAR, FN_ATAN2, FN_LOG,
    FN_BESSELJ, FN_BESSELY, FN_BESSELI, FN_BESSELK, FN_POLYGAMMA, FN_BETA, FN_ZETADERIV,
    FN_HERMITE, FN_ASSOCLEGENDRE, FN_SPHERICALHARMONIC, FN_COUNT
};

/* helpers.rs:9-11 */
static double eval_sinc(double x) { return fabs(x) < SAF_EPSILON ? 1.0 : sin(x) / x; }

/* helpers.rs:15-25 */
static int round_to_i32(double x, int32_t *out) {
    double rounded = round(x);
    if (!(rounded >= -2147483648.0 && rounded <= 2147483647.0)) return 0;
    *out = (int32_t)rounded;
    return 1;
}

double saf_oracle_powi(double v, int32_t n) { return saf_o_powi(v, n); }

/* builtins.rs:38-86 */
double saf_oracle_builtin1(uint32_t fnop, double x) {
    switch (fnop & 0xffu) { /* transmute_copy reads the low byte, macros.rs:311 */
    case FN_TAN: return tan(x);
    case FN_COT: return 1.0 / tan(x);
    case FN_SEC: return 1.0 / cos(x);
    case FN_CSC: return 1.0 / sin(x);
    case FN_ASIN: return asin(x);
    case FN_ACOS: return acos(x);
    case FN_ATAN: return atan(x);
    case FN_ACOT: return SAF_FRAC_PI_2 - atan(x);
    case FN_ASEC: return acos(1.0 / x);
    case FN_ACSC: return asin(1.0 / x);
    case FN_SINH: return sinh(x);
    case FN_COSH: return cosh(x);
    case FN_TANH: return tanh(x);
    case FN_COTH: return 1.0 / tanh(x);
    case FN_SECH: return 1.0 / cosh(x);
    case FN_CSCH: return 1.0 / sinh(x);
    case FN_ASINH: return saf_o_asinh(x);
    case FN_ACOSH: return saf_o_acosh(x);
    case FN_ATANH: return saf_o_atanh(x);
    case FN_ACOTH: return saf_o_atanh(1.0 / x);
    case FN_ACSCH: return saf_o_asinh(1.0 / x);
    case FN_ASECH: return saf_o_acosh(1.0 / x);
    case FN_EXPM1: return expm1(x);
    case FN_EXPNEG: return exp(-x);
    case FN_LOG1P: return log1p(x);
    case FN_CBRT: return cbrt(x);
    case FN_ABS: return fabs(x);
    case FN_SIGNUM: return saf_o_signum(x);
    case FN_FLOOR: return floor(x);
    case FN_CEIL: return ceil(x);
    case FN_ROUND: return round(x);
    case FN_ERF: return saf_o_erf(x);
    case FN_ERFC: return saf_o_erfc(x);
    case FN_GAMMA: return saf_o_gamma(x);
    case FN_LGAMMA: return saf_o_lgamma(x);
    case FN_DIGAMMA: return saf_o_digamma(x);
    case FN_TRIGAMMA: return saf_o_trigamma(x);
    case FN_TETRAGAMMA: return saf_o_tetragamma(x);
    case FN_SINC: return eval_sinc(x);
    case FN_LAMBERTW: return saf_o_lambert_w(x);
    case FN_ELLIPTICK: return saf_o_elliptic_k(x);
    case FN_ELLIPTICE: return saf_o_elliptic_e(x);
    case FN_ZETA: return saf_o_zeta(x);
    case FN_EXPPOLAR: return exp(x); /* polar.rs:8-10 */
    default: return NAN; /* unreachable_builtin, builtins.rs:21-26 (release build) */
    }
}

/* builtins.rs:90-115 */
double saf_oracle_builtin2(uint32_t fnop, double x1, double x2) {
    int32_t n;
    switch (fnop & 0xffu) {
    case FN_ATAN2: return atan2(x1, x2);
    case FN_LOG:
        if (x1 <= 0.0 || x1 == 1.0 || x2 < 0.0) return NAN;
        return log(x2) / log(x1); /* f64::log(self, base) = self.ln() / base.ln() */
    case FN_BESSELJ: return round_to_i32(x1, &n) ? saf_o_bessel_j(n, x2) : NAN;
    case FN_BESSELY: return round_to_i32(x1, &n) ? saf_o_bessel_y(n, x2) : NAN;
    case FN_BESSELI: return round_to_i32(x1, &n) ? saf_o_bessel_i(n, x2) : NAN;
    case FN_BESSELK: return round_to_i32(x1, &n) ? saf_o_bessel_k(n, x2) : NAN;
    case FN_POLYGAMMA: return round_to_i32(x1, &n) ? saf_o_polygamma(n, x2) : NAN;
    case FN_BETA: return saf_o_beta(x1, x2);
    case FN_ZETADERIV: return round_to_i32(x1, &n) ? saf_o_zeta_deriv(n, x2) : NAN;
    case FN_HERMITE: return round_to_i32(x1, &n) ? saf_o_hermite(n, x2) : NAN;
    default: return NAN;
    }
}

/* builtins.rs:119-127 */
double saf_oracle_builtin3(uint32_t fnop, double x1, double x2, double x3) {
    int32_t l, m;
    if ((fnop & 0xffu) == FN_ASSOCLEGENDRE) {
        if (round_to_i32(x1, &l) && round_to_i32(x2, &m)) return saf_o_assoc_legendre(l, m, x3);
        return NAN;
    }
    return NAN;
}

/* builtins.rs:131-139 */
double saf_oracle_builtin4(uint32_t fnop, double x1, double x2, double x3, double x4) {
    int32_t l, m;
    if ((fnop & 0xffu) == FN_SPHERICALHARMONIC) {
        if (round_to_i32(x1, &l) && round_to_i32(x2, &m))
            return saf_o_spherical_harmonic(l, m, x3, x4);
        return NAN;
    }
    return NAN;
}

/* macros.rs:1-357, scalar arms of :359-411.  Sources are read before the
 * destination is written (in-place instructions exist, fusion.rs:742-753). */
void saf_oracle_exec(const saf_oracle_program *p, double *regs) {
    const uint32_t *pc = p->flat;
    const uint32_t *pool = p->arg_pool;
    for (;;) {
        uint32_t opcode = *pc++;
        switch (opcode) {
        case 0: return; /* End */
        case 1: { uint32_t d = pc[0], s = pc[1]; pc += 2; regs[d] = regs[s]; } break;
        case 2: { uint32_t d = pc[0], s = pc[1]; pc += 2; regs[d] = -regs[s]; } break;
        case 3: { /* SinCos: f64::sin_cos = (sin, cos) */
            uint32_t sd = pc[0], cd = pc[1], a = pc[2]; pc += 3;
            double v = regs[a];
            double s = sin(v), c = cos(v);
            regs[sd] = s; regs[cd] = c;
        } break;
        case 4: { uint32_t d = pc[0], a = pc[1], b = pc[2]; pc += 3; regs[d] = regs[a] + regs[b]; } break;
        case 5: { uint32_t d = pc[0], a = pc[1], b = pc[2], c = pc[3]; pc += 4;
                  regs[d] = regs[a] + regs[b] + regs[c]; } break;
        case 6: { uint32_t d = pc[0], a = pc[1], b = pc[2], c = pc[3], e = pc[4]; pc += 5;
                  regs[d] = regs[a] + regs[b] + regs[c] + regs[e]; } break;
        case 7: { /* AddN */
            uint32_t d = pc[0], start = pc[1], count = pc[2]; pc += 3;
            double sum = regs[pool[start]];
            for (uint32_t i = 1; i < count; ++i) sum += regs[pool[start + i]];
            regs[d] = sum;
        } break;
        case 8: { uint32_t d = pc[0], a = pc[1], b = pc[2]; pc += 3; regs[d] = regs[a] * regs[b]; } break;
        case 9: { uint32_t d = pc[0], a = pc[1], b = pc[2], c = pc[3]; pc += 4;
                  regs[d] = regs[a] * regs[b] * regs[c]; } break;
        case 10: { uint32_t d = pc[0], a = pc[1], b = pc[2], c = pc[3], e = pc[4]; pc += 5;
                   regs[d] = regs[a] * regs[b] * regs[c] * regs[e]; } break;
        case 11: { /* MulN */
            uint32_t d = pc[0], start = pc[1], count = pc[2]; pc += 3;
            double prod = regs[pool[start]];
            for (uint32_t i = 1; i < count; ++i) prod *= regs[pool[start + i]];
            regs[d] = prod;
        } break;
        case 12: { uint32_t d = pc[0], a = pc[1], b = pc[2]; pc += 3; regs[d] = regs[a] - regs[b]; } break;
        case 13: { uint32_t d = pc[0], a = pc[1], b = pc[2]; pc += 3; regs[d] = regs[a] / regs[b]; } break;
        case 14: { uint32_t d = pc[0], a = pc[1], b = pc[2]; pc += 3; regs[d] = pow(regs[a], regs[b]); } break;
        case 15: { uint32_t d = pc[0], a = pc[1], b = pc[2], c = pc[3]; pc += 4;
                   regs[d] = fma(regs[a], regs[b], regs[c]); } break;
        case 16: { uint32_t d = pc[0], a = pc[1], b = pc[2], c = pc[3]; pc += 4;
                   regs[d] = fma(regs[a], regs[b], -regs[c]); } break;
        case 17: { uint32_t d = pc[0], a = pc[1], b = pc[2]; pc += 3; regs[d] = -(regs[a] * regs[b]); } break;
        case 18: { uint32_t d = pc[0], a = pc[1], b = pc[2], c = pc[3]; pc += 4;
                   regs[d] = fma(-regs[a], regs[b], regs[c]); } break;
        case 19: { uint32_t d = pc[0], a = pc[1], b = pc[2], c = pc[3]; pc += 4;
                   regs[d] = fma(-regs[a], regs[b], -regs[c]); } break;
        case 20: { uint32_t d = pc[0], s = pc[1]; pc += 2; double v = regs[s]; regs[d] = v * v; } break;
        case 21: { uint32_t d = pc[0], s = pc[1]; pc += 2; double v = regs[s]; regs[d] = v * v * v; } break;
        case 22: { uint32_t d = pc[0], s = pc[1]; pc += 2; double v = regs[s]; double v2 = v * v; regs[d] = v2 * v2; } break;
        case 23: { uint32_t d = pc[0], s = pc[1]; pc += 2; double v = regs[s]; regs[d] = v * sqrt(v); } break;
        case 24: { uint32_t d = pc[0], s = pc[1]; pc += 2; double v = regs[s]; regs[d] = 1.0 / (v * sqrt(v)); } break;
        case 25: { uint32_t d = pc[0], s = pc[1]; pc += 2; regs[d] = 1.0 / sqrt(regs[s]); } break;
        case 26: { uint32_t d = pc[0], s = pc[1]; pc += 2; double v = regs[s]; regs[d] = 1.0 / (v * v); } break;
        case 27: { uint32_t d = pc[0], s = pc[1]; pc += 2; double v = regs[s]; regs[d] = 1.0 / (v * v * v); } break;
        case 28: { uint32_t d = pc[0], s = pc[1]; pc += 2; regs[d] = 1.0 / regs[s]; } break;
        case 29: { uint32_t d = pc[0], s = pc[1]; int32_t n = (int32_t)pc[2]; pc += 3;
                   regs[d] = saf_o_powi(regs[s], n); } break;
        case 30: { uint32_t d = pc[0], a = pc[1]; pc += 2; regs[d] = sin(regs[a]); } break;
        case 31: { uint32_t d = pc[0], a = pc[1]; pc += 2; regs[d] = cos(regs[a]); } break;
        case 32: { uint32_t d = pc[0], a = pc[1]; pc += 2; regs[d] = exp(regs[a]); } break;
        case 33: { uint32_t d = pc[0], a = pc[1]; pc += 2; regs[d] = log(regs[a]); } break;
        case 34: { uint32_t d = pc[0], a = pc[1]; pc += 2; regs[d] = sqrt(regs[a]); } break;
        case 35: { uint32_t d = pc[0], s = pc[1]; pc += 2; regs[d] = 1.0 / expm1(regs[s]); } break;
        case 36: { uint32_t d = pc[0], s = pc[1]; pc += 2; double v = regs[s]; regs[d] = exp(v * v); } break;
        case 37: { uint32_t d = pc[0], s = pc[1]; pc += 2; double v = regs[s]; regs[d] = exp(-(v * v)); } break;
        case 38: { uint32_t d = pc[0], op = pc[1], a = pc[2]; pc += 3;
                   regs[d] = saf_oracle_builtin1(op, regs[a]); } break;
        case 39: { uint32_t d = pc[0], op = pc[1], a1 = pc[2], a2 = pc[3]; pc += 4;
                   regs[d] = saf_oracle_builtin2(op, regs[a1], regs[a2]); } break;
        case 40: { uint32_t d = pc[0], op = pc[1], a1 = pc[2], a2 = pc[3], a3 = pc[4]; pc += 5;
                   regs[d] = saf_oracle_builtin3(op, regs[a1], regs[a2], regs[a3]); } break;
        case 41: { uint32_t d = pc[0], op = pc[1], a1 = pc[2], a2 = pc[3], a3 = pc[4], a4 = pc[5]; pc += 6;
                   regs[d] = saf_oracle_builtin4(op, regs[a1], regs[a2], regs[a3], regs[a4]); } break;
        default:
            /* unreachable_unchecked in the reference (macros.rs:355); the
             * validator rejects such programs before they get here. */
            return;
        }
    }
}

/* number of operand words after each opcode, assemble.rs:9-103 */
static const int8_t OPERAND_WORDS[42] = {
    0, 2, 2, 3, 3, 4, 5, 3, 3, 4, 5, 3, 3, 3, 3, 4, 4, 3, 4, 4, 2, 2,
    2, 2, 2, 2, 2, 2, 2, 3, 2, 2, 2, 2, 2, 2, 2, 2, 3, 4, 5, 6};

static void set_err(char *err, size_t errlen, const char *msg, long a, long b) {
    if (err && errlen) snprintf(err, errlen, msg, a, b);
}

/* helper.rs:98-196 semantics on the flat stream: bounds, read-before-definition,
 * arg-pool ranges; plus stream well-formedness (known opcode, no truncation). */
int saf_oracle_validate(const saf_oracle_program *p, char *err, size_t errlen) {
    uint64_t ws = p->workspace_size;
    uint64_t const_limit = (uint64_t)p->param_count + p->n_consts;
    if (const_limit > ws) {
        set_err(err, errlen, "invalid register layout during validation (%ld > %ld)", (long)const_limit, (long)ws);
        return SAF_ORACLE_ERR_INVALID_PROGRAM;
    }
    unsigned char *defined = (unsigned char *)calloc(ws ? ws : 1, 1);
    if (!defined) return SAF_ORACLE_ERR_ALLOC;
    memset(defined, 1, const_limit);
    uint64_t pc = 0;
    long instr_idx = 0;
    int rc = SAF_ORACLE_OK;
#define CHECK_READ(r)                                                                              \
    do {                                                                                           \
        uint32_t _r = (r);                                                                         \
        if (_r >= ws) { set_err(err, errlen, "source register R%ld out of bounds at instruction %ld", (long)_r, instr_idx); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; } \
        if (!defined[_r]) { set_err(err, errlen, "source register R%ld read before definition at instruction %ld", (long)_r, instr_idx); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; } \
    } while (0)
#define CHECK_WRITE(r)                                                                             \
    do {                                                                                           \
        uint32_t _r = (r);                                                                         \
        if (_r >= ws) { set_err(err, errlen, "destination register R%ld out of bounds at instruction %ld", (long)_r, instr_idx); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; } \
        defined[_r] = 1;                                                                           \
    } while (0)
    for (;; ++instr_idx) {
        if (pc >= p->n_words) { set_err(err, errlen, "bytecode stream not terminated by End (pc %ld of %ld)", (long)pc, (long)p->n_words); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; }
        uint32_t op = p->flat[pc];
        if (op == 0) break;
        if (op > 41) { set_err(err, errlen, "unknown opcode %ld at instruction %ld", (long)op, instr_idx); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; }
        int nw = OPERAND_WORDS[op];
        if (pc + 1 + (uint64_t)nw > p->n_words) { set_err(err, errlen, "truncated instruction %ld (opcode %ld)", instr_idx, (long)op); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; }
        const uint32_t *w = p->flat + pc + 1;
        switch (op) {
        case 3: CHECK_READ(w[2]); CHECK_WRITE(w[0]); CHECK_WRITE(w[1]); break;
        case 7: case 11: {
            uint64_t start = w[1], count = w[2];
            if (start + count > p->n_pool) { set_err(err, errlen, "arg pool slice out of bounds at instruction %ld (end %ld)", instr_idx, (long)(start + count)); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; }
            if (count == 0) { set_err(err, errlen, "empty arg pool slice at instruction %ld%.0ld", instr_idx, 0L); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; }
            for (uint64_t i = 0; i < count; ++i) CHECK_READ(p->arg_pool[start + i]);
            CHECK_WRITE(w[0]);
        } break;
        case 29: CHECK_READ(w[1]); CHECK_WRITE(w[0]); break; /* Powi: w[2] is an immediate */
        case 38: case 39: case 40: case 41: { /* w[1] is the FnOp */
            for (int i = 2; i < nw; ++i) CHECK_READ(w[i]);
            if ((w[1] & 0xffu) >= FN_COUNT) { set_err(err, errlen, "unknown FnOp %ld at instruction %ld", (long)(w[1] & 0xffu), instr_idx); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; }
            CHECK_WRITE(w[0]);
        } break;
        default:
            for (int i = 1; i < nw; ++i) CHECK_READ(w[i]);
            CHECK_WRITE(w[0]);
        }
        pc += 1 + (uint64_t)nw;
    }
    for (uint32_t k = 0; k < p->n_results; ++k) {
        uint32_t r = p->result_regs[k];
        if (r >= ws || !defined[r]) { set_err(err, errlen, "result register R%ld undefined or out of bounds (result %ld)", (long)r, (long)k); rc = SAF_ORACLE_ERR_INVALID_PROGRAM; goto done; }
    }
done:
    free(defined);
    return rc;
#undef CHECK_READ
#undef CHECK_WRITE
}

/* scalar.rs:93-108 */
static void setup_registers(const saf_oracle_program *p, const double *params, size_t n_params, double *regs) {
    size_t pc = p->param_count < n_params ? p->param_count : n_params;
    if (pc) memcpy(regs, params, pc * sizeof(double));
    for (size_t i = pc; i < p->param_count; ++i) regs[i] = 0.0;
    if (p->n_consts) memcpy(regs + p->param_count, p->consts, p->n_consts * sizeof(double));
}

int saf_oracle_evaluate(const saf_oracle_program *p, const double *params, size_t n_params, double *out) {
    double *regs = (double *)malloc(sizeof(double) * (p->workspace_size ? p->workspace_size : 1));
    if (!regs) return SAF_ORACLE_ERR_ALLOC;
    setup_registers(p, params, n_params, regs);
    saf_oracle_exec(p, regs);
    for (uint32_t k = 0; k < p->n_results; ++k) out[k] = regs[p->result_regs[k]];
    free(regs);
    return SAF_ORACLE_OK;
}

/* scalar.rs:172-230 on the point range [begin, end) with a caller workspace */
static void eval_range_scalar(const saf_oracle_program *p, const double *const *cols,
                              const size_t *col_lens, size_t n_cols, double *const *outs,
                              size_t begin, size_t end, double *regs) {
    if (p->n_consts) memcpy(regs + p->param_count, p->consts, p->n_consts * sizeof(double));
    size_t provided = p->param_count < n_cols ? p->param_count : n_cols;
    for (size_t c = provided; c < p->param_count; ++c) regs[c] = 0.0;
    for (size_t i = begin; i < end; ++i) {
        for (size_t c = 0; c < provided; ++c) {
            size_t len = col_lens[c];
            regs[c] = (i < len) ? cols[c][i] : cols[c][len - 1];
        }
        saf_oracle_exec(p, regs);
        for (uint32_t k = 0; k < p->n_results; ++k) outs[k][i] = regs[p->result_regs[k]];
    }
}

int saf_oracle_eval_batch(const saf_oracle_program *p, const double *const *cols,
                          const size_t *col_lens, size_t n_cols, double *const *outs, size_t n_points) {
    if (n_points == 0) return SAF_ORACLE_OK; /* scalar.rs:157-159 */
    double *regs = (double *)malloc(sizeof(double) * (p->workspace_size ? p->workspace_size : 1));
    if (!regs) return SAF_ORACLE_ERR_ALLOC;
    eval_range_scalar(p, cols, col_lens, n_cols, outs, 0, n_points, regs);
    free(regs);
    return SAF_ORACLE_OK;
}

int saf_oracle_online_cores(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    const saf_oracle_program *p;
    const double *const *cols;
    const size_t *col_lens;
    size_t n_cols;
    double *const *outs;
    size_t n_points;
    atomic_size_t next_chunk;
    size_t n_chunks;
    int failed;
} chunk_job;

static void *chunk_worker(void *arg) {
    chunk_job *job = (chunk_job *)arg;
    const saf_oracle_program *p = job->p;
    /* one workspace per worker, batch.rs:59-64 */
    double *regs = (double *)malloc(sizeof(double) * (p->workspace_size ? p->workspace_size : 1));
    if (!regs) { job->failed = 1; return NULL; }
    for (;;) {
        size_t c = atomic_fetch_add(&job->next_chunk, 1);
        if (c >= job->n_chunks) break;
        size_t start = c * SAF_CHUNK_SIZE;
        size_t end = start + SAF_CHUNK_SIZE;
        if (end > job->n_points) end = job->n_points;
        eval_range_scalar(p, job->cols, job->col_lens, job->n_cols, job->outs, start, end, regs);
    }
    free(regs);
    return NULL;
}

static int run_threads(chunk_job *job, int n_threads) {
    if (n_threads <= 0) n_threads = saf_oracle_online_cores();
    if ((size_t)n_threads > job->n_chunks) n_threads = (int)job->n_chunks;
    if (n_threads <= 1) {
        chunk_worker(job);
        return job->failed ? SAF_ORACLE_ERR_ALLOC : SAF_ORACLE_OK;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    if (!th) return SAF_ORACLE_ERR_ALLOC;
    int started = 0;
    for (int t = 0; t < n_threads; ++t) {
        if (pthread_create(&th[t], NULL, chunk_worker, job) != 0) break;
        ++started;
    }
    if (started == 0) chunk_worker(job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    return job->failed ? SAF_ORACLE_ERR_ALLOC : SAF_ORACLE_OK;
}

/* batch.rs:37-78 */
int saf_oracle_run_chunked(const saf_oracle_program *p, const double *const *cols,
                           const size_t *col_lens, size_t n_cols, double *const *outs,
                           size_t n_points, int n_threads) {
    if (n_cols == 0) {
        if (n_points == 0) return SAF_ORACLE_OK;
    } else {
        for (size_t c = 0; c < n_cols; ++c)
            if (col_lens[c] != n_points) return SAF_ORACLE_ERR_COLUMN_LENGTH_MISMATCH;
    }
    if (n_points < SAF_CHUNK_SIZE) return saf_oracle_eval_batch(p, cols, col_lens, n_cols, outs, n_points);
    chunk_job job;
    memset(&job, 0, sizeof job);
    job.p = p; job.cols = cols; job.col_lens = col_lens; job.n_cols = n_cols; job.outs = outs;
    job.n_points = n_points;
    atomic_init(&job.next_chunk, 0);
    job.n_chunks = (n_points + SAF_CHUNK_SIZE - 1) / SAF_CHUNK_SIZE;
    return run_threads(&job, n_threads);
}

typedef struct {
    const saf_oracle_program *p;
    double *const *state;
    size_t n_state, n_points;
    uint64_t iters;
    atomic_size_t next_chunk;
    size_t n_chunks;
    int failed;
} iter_job;

static void *iter_worker(void *arg) {
    iter_job *job = (iter_job *)arg;
    const saf_oracle_program *p = job->p;
    double *regs = (double *)malloc(sizeof(double) * (p->workspace_size ? p->workspace_size : 1));
    double *tmp = (double *)malloc(sizeof(double) * (job->n_state ? job->n_state : 1));
    if (!regs || !tmp) { job->failed = 1; free(regs); free(tmp); return NULL; }
    if (p->n_consts) memcpy(regs + p->param_count, p->consts, p->n_consts * sizeof(double));
    size_t provided = p->param_count < job->n_state ? p->param_count : job->n_state;
    for (;;) {
        size_t c = atomic_fetch_add(&job->next_chunk, 1);
        if (c >= job->n_chunks) break;
        size_t start = c * SAF_CHUNK_SIZE, end = start + SAF_CHUNK_SIZE;
        if (end > job->n_points) end = job->n_points;
        /* points are independent, so iterating each point `iters` times in
         * place is the same computation as `iters` whole-array passes. */
        for (size_t i = start; i < end; ++i) {
            for (size_t j = 0; j < job->n_state; ++j) tmp[j] = job->state[j][i];
            for (uint64_t it = 0; it < job->iters; ++it) {
                for (size_t j = 0; j < provided; ++j) regs[j] = tmp[j];
                for (size_t j = provided; j < p->param_count; ++j) regs[j] = 0.0;
                saf_oracle_exec(p, regs);
                for (size_t j = 0; j < job->n_state; ++j) tmp[j] = regs[p->result_regs[j]];
            }
            for (size_t j = 0; j < job->n_state; ++j) job->state[j][i] = tmp[j];
        }
    }
    free(regs);
    free(tmp);
    return NULL;
}

int saf_oracle_iterate(const saf_oracle_program *p, double *const *state, size_t n_state,
                       size_t n_points, uint64_t iters, int n_threads) {
    if (p->n_results != n_state) return SAF_ORACLE_ERR_INVALID_PROGRAM;
    if (n_points == 0 || iters == 0) return SAF_ORACLE_OK;
    iter_job job;
    memset(&job, 0, sizeof job);
    job.p = p; job.state = state; job.n_state = n_state; job.n_points = n_points; job.iters = iters;
    atomic_init(&job.next_chunk, 0);
    job.n_chunks = (n_points + SAF_CHUNK_SIZE - 1) / SAF_CHUNK_SIZE;
    if (n_threads <= 0) n_threads = saf_oracle_online_cores();
    if ((size_t)n_threads > job.n_chunks) n_threads = (int)job.n_chunks;
    if (n_threads <= 1) { iter_worker(&job); return job.failed ? SAF_ORACLE_ERR_ALLOC : SAF_ORACLE_OK; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    if (!th) return SAF_ORACLE_ERR_ALLOC;
    int started = 0;
    for (int t = 0; t < n_threads; ++t) {
        if (pthread_create(&th[t], NULL, iter_worker, &job) != 0) break;
        ++started;
    }
    if (started == 0) iter_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    return job.failed ? SAF_ORACLE_ERR_ALLOC : SAF_ORACLE_OK;
}
