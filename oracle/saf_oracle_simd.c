/*
 * saf_oracle_simd.c -- the reference's f64x4 path, restated.  TEST INFRASTRUCTURE ONLY (see saf_oracle.h).
 *
 * What it follows (paths relative to /root/reference/src/evaluator/logic/bytecode/execute):
 *   - eval_batch_simd                 engine/simd.rs:53-114   constants splatted once, missing columns 0.0, groups of
 *                                                             four points, a column shorter than i + 4 is splatted with
 *                                                             its LAST value for the whole group (:76-80), scalar tail
 *   - the simd arms of dispatch_loop! engine/macros.rs:1-357 with :359-411: + - * / sqrt abs neg mul_add are lane-wise
 *                                                             vector operations, every transcendental is four scalar
 *                                                             std calls (to_array().map(..))
 *   - eval_builtin*_simd              engine/builtins.rs:148-299: Abs is a vector op, the rest lane by lane; the
 *                                                             reciprocal forms divide as vectors
 *   - run_chunked_evaluator           drivers/batch.rs:37-78  256-point chunks, one f64x4 workspace per worker
 *
 * Third-party arithmetic not under /root/reference: `wide` 1.3.0 f64x4 (Cargo.toml:42) -- lane-wise IEEE operations;
 * its mul_add is a true FMA only when the crate is built with the `fma` target feature (SURVEY.md section 2 quirk 3,
 * "parity unpinned").  This restatement uses a fused multiply-add (what the scalar path's f64::mul_add always is), so
 * that the f64x4 body and the scalar oracle agree bit for bit and either can check the GPU.
 *
 * Needs AVX2 + FMA at run time (checked by saf_oracle_has_simd); the file is compiled with per-function target
 * attributes so that the rest of the oracle stays baseline x86-64.
 */
#define _GNU_SOURCE
#include "saf_oracle.h"
#include "saf_oracle_math.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define SAF_SIMD_TARGET __attribute__((target("avx2,fma")))

typedef __m256d v4;

#define SAF_CHUNK_SIZE 256 /* batch.rs:9 */

int saf_oracle_has_simd(void) { return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma"); }

#define LANES(expr_of_x)                                                                           \
    ({                                                                                             \
        double a_[4] __attribute__((aligned(32)));                                                 \
        _mm256_store_pd(a_, x_);                                                                   \
        for (int l_ = 0; l_ < 4; ++l_) { double x = a_[l_]; a_[l_] = (expr_of_x); }                \
        _mm256_load_pd(a_);                                                                        \
    })

SAF_SIMD_TARGET static inline v4 map1(v4 x_, double (*f)(double)) { return LANES(f(x)); }
SAF_SIMD_TARGET static inline v4 splat(double v) { return _mm256_set1_pd(v); }
SAF_SIMD_TARGET static inline v4 vneg(v4 v) { return _mm256_xor_pd(v, _mm256_set1_pd(-0.0)); }

/* builtins.rs:148-201 (one argument) */
SAF_SIMD_TARGET static v4 builtin1_simd(uint32_t fnop, v4 x_) {
    const v4 one = splat(1.0);
    switch (fnop & 0xffu) {
    case 31: /* Abs: a vector operation (builtins.rs:149-151) */ return _mm256_andnot_pd(_mm256_set1_pd(-0.0), x_);
    case 3: /* Cot */ return _mm256_div_pd(one, map1(x_, tan));
    case 4: /* Sec */ return _mm256_div_pd(one, map1(x_, cos));
    case 5: /* Csc */ return _mm256_div_pd(one, map1(x_, sin));
    case 9: /* Acot */ return _mm256_sub_pd(splat(1.57079632679489661923132169163975144), map1(x_, atan));
    case 15: /* Coth */ return _mm256_div_pd(one, map1(x_, tanh));
    case 16: /* Sech */ return _mm256_div_pd(one, map1(x_, cosh));
    case 17: /* Csch */ return _mm256_div_pd(one, map1(x_, sinh));
    default: /* lane by lane; 1/x inside Asec.. is the same IEEE division either way */
        return LANES(saf_oracle_builtin1(fnop, x));
    }
}

/* macros.rs:1-357, simd arms */
SAF_SIMD_TARGET static void exec_simd(const saf_oracle_program *p, v4 *regs) {
    const uint32_t *pc = p->flat;
    const uint32_t *pool = p->arg_pool;
    const v4 one = splat(1.0);
    for (;;) {
        const uint32_t opcode = *pc++;
        switch (opcode) {
        case 0: return;
        case 1: regs[pc[0]] = regs[pc[1]]; pc += 2; break;
        case 2: regs[pc[0]] = vneg(regs[pc[1]]); pc += 2; break;
        case 3: {
            double a[4] __attribute__((aligned(32))), s[4] __attribute__((aligned(32))), c[4] __attribute__((aligned(32)));
            _mm256_store_pd(a, regs[pc[2]]);
            for (int l = 0; l < 4; ++l) { s[l] = sin(a[l]); c[l] = cos(a[l]); }
            regs[pc[0]] = _mm256_load_pd(s);
            regs[pc[1]] = _mm256_load_pd(c);
            pc += 3;
        } break;
        case 4: regs[pc[0]] = _mm256_add_pd(regs[pc[1]], regs[pc[2]]); pc += 3; break;
        case 5: regs[pc[0]] = _mm256_add_pd(_mm256_add_pd(regs[pc[1]], regs[pc[2]]), regs[pc[3]]); pc += 4; break;
        case 6:
            regs[pc[0]] = _mm256_add_pd(_mm256_add_pd(_mm256_add_pd(regs[pc[1]], regs[pc[2]]), regs[pc[3]]), regs[pc[4]]);
            pc += 5;
            break;
        case 7: {
            const uint32_t d = pc[0], start = pc[1], count = pc[2];
            pc += 3;
            v4 sum = regs[pool[start]];
            for (uint32_t i = 1; i < count; ++i) sum = _mm256_add_pd(sum, regs[pool[start + i]]);
            regs[d] = sum;
        } break;
        case 8: regs[pc[0]] = _mm256_mul_pd(regs[pc[1]], regs[pc[2]]); pc += 3; break;
        case 9: regs[pc[0]] = _mm256_mul_pd(_mm256_mul_pd(regs[pc[1]], regs[pc[2]]), regs[pc[3]]); pc += 4; break;
        case 10:
            regs[pc[0]] = _mm256_mul_pd(_mm256_mul_pd(_mm256_mul_pd(regs[pc[1]], regs[pc[2]]), regs[pc[3]]), regs[pc[4]]);
            pc += 5;
            break;
        case 11: {
            const uint32_t d = pc[0], start = pc[1], count = pc[2];
            pc += 3;
            v4 prod = regs[pool[start]];
            for (uint32_t i = 1; i < count; ++i) prod = _mm256_mul_pd(prod, regs[pool[start + i]]);
            regs[d] = prod;
        } break;
        case 12: regs[pc[0]] = _mm256_sub_pd(regs[pc[1]], regs[pc[2]]); pc += 3; break;
        case 13: regs[pc[0]] = _mm256_div_pd(regs[pc[1]], regs[pc[2]]); pc += 3; break;
        case 14: {
            double b[4] __attribute__((aligned(32))), e[4] __attribute__((aligned(32)));
            _mm256_store_pd(b, regs[pc[1]]);
            _mm256_store_pd(e, regs[pc[2]]);
            for (int l = 0; l < 4; ++l) b[l] = pow(b[l], e[l]);
            regs[pc[0]] = _mm256_load_pd(b);
            pc += 3;
        } break;
        case 15: regs[pc[0]] = _mm256_fmadd_pd(regs[pc[1]], regs[pc[2]], regs[pc[3]]); pc += 4; break;
        case 16: regs[pc[0]] = _mm256_fmsub_pd(regs[pc[1]], regs[pc[2]], regs[pc[3]]); pc += 4; break;
        case 17: regs[pc[0]] = vneg(_mm256_mul_pd(regs[pc[1]], regs[pc[2]])); pc += 3; break;
        case 18: regs[pc[0]] = _mm256_fnmadd_pd(regs[pc[1]], regs[pc[2]], regs[pc[3]]); pc += 4; break;
        case 19: regs[pc[0]] = _mm256_fnmsub_pd(regs[pc[1]], regs[pc[2]], regs[pc[3]]); pc += 4; break;
        case 20: { const v4 v = regs[pc[1]]; regs[pc[0]] = _mm256_mul_pd(v, v); pc += 2; } break;
        case 21: { const v4 v = regs[pc[1]]; regs[pc[0]] = _mm256_mul_pd(_mm256_mul_pd(v, v), v); pc += 2; } break;
        case 22: { const v4 v = regs[pc[1]]; const v4 v2 = _mm256_mul_pd(v, v); regs[pc[0]] = _mm256_mul_pd(v2, v2); pc += 2; } break;
        case 23: { const v4 v = regs[pc[1]]; regs[pc[0]] = _mm256_mul_pd(v, _mm256_sqrt_pd(v)); pc += 2; } break;
        case 24: { const v4 v = regs[pc[1]]; regs[pc[0]] = _mm256_div_pd(one, _mm256_mul_pd(v, _mm256_sqrt_pd(v))); pc += 2; } break;
        case 25: regs[pc[0]] = _mm256_div_pd(one, _mm256_sqrt_pd(regs[pc[1]])); pc += 2; break;
        case 26: { const v4 v = regs[pc[1]]; regs[pc[0]] = _mm256_div_pd(one, _mm256_mul_pd(v, v)); pc += 2; } break;
        case 27: { const v4 v = regs[pc[1]]; regs[pc[0]] = _mm256_div_pd(one, _mm256_mul_pd(_mm256_mul_pd(v, v), v)); pc += 2; } break;
        case 28: regs[pc[0]] = _mm256_div_pd(one, regs[pc[1]]); pc += 2; break;
        case 29: {
            const int32_t n = (int32_t)pc[2];
            const v4 x_ = regs[pc[1]];
            regs[pc[0]] = LANES(saf_o_powi(x, n));
            pc += 3;
        } break;
        case 30: regs[pc[0]] = map1(regs[pc[1]], sin); pc += 2; break;
        case 31: regs[pc[0]] = map1(regs[pc[1]], cos); pc += 2; break;
        case 32: regs[pc[0]] = map1(regs[pc[1]], exp); pc += 2; break;
        case 33: regs[pc[0]] = map1(regs[pc[1]], log); pc += 2; break;
        case 34: regs[pc[0]] = _mm256_sqrt_pd(regs[pc[1]]); pc += 2; break;
        case 35: regs[pc[0]] = _mm256_div_pd(one, map1(regs[pc[1]], expm1)); pc += 2; break;
        case 36: { const v4 v = regs[pc[1]]; regs[pc[0]] = map1(_mm256_mul_pd(v, v), exp); pc += 2; } break;
        case 37: { const v4 v = regs[pc[1]]; regs[pc[0]] = map1(vneg(_mm256_mul_pd(v, v)), exp); pc += 2; } break;
        case 38: regs[pc[0]] = builtin1_simd(pc[1], regs[pc[2]]); pc += 3; break;
        case 39: {
            double a[4] __attribute__((aligned(32))), b[4] __attribute__((aligned(32)));
            _mm256_store_pd(a, regs[pc[2]]);
            _mm256_store_pd(b, regs[pc[3]]);
            for (int l = 0; l < 4; ++l) a[l] = saf_oracle_builtin2(pc[1], a[l], b[l]);
            regs[pc[0]] = _mm256_load_pd(a);
            pc += 4;
        } break;
        case 40: {
            double a[4] __attribute__((aligned(32))), b[4] __attribute__((aligned(32))), c[4] __attribute__((aligned(32)));
            _mm256_store_pd(a, regs[pc[2]]);
            _mm256_store_pd(b, regs[pc[3]]);
            _mm256_store_pd(c, regs[pc[4]]);
            for (int l = 0; l < 4; ++l) a[l] = saf_oracle_builtin3(pc[1], a[l], b[l], c[l]);
            regs[pc[0]] = _mm256_load_pd(a);
            pc += 5;
        } break;
        case 41: {
            double a[4] __attribute__((aligned(32))), b[4] __attribute__((aligned(32))), c[4] __attribute__((aligned(32))),
                d[4] __attribute__((aligned(32)));
            _mm256_store_pd(a, regs[pc[2]]);
            _mm256_store_pd(b, regs[pc[3]]);
            _mm256_store_pd(c, regs[pc[4]]);
            _mm256_store_pd(d, regs[pc[5]]);
            for (int l = 0; l < 4; ++l) a[l] = saf_oracle_builtin4(pc[1], a[l], b[l], c[l], d[l]);
            regs[pc[0]] = _mm256_load_pd(a);
            pc += 6;
        } break;
        default: return; /* rejected by the validator */
        }
    }
}

/* simd.rs:53-114 on the point range [begin, end) of columns that are indexed from `begin` on */
SAF_SIMD_TARGET static void eval_range_simd(const saf_oracle_program *p, const double *const *cols, const size_t *col_lens,
                                            size_t n_cols, double *const *outs, size_t begin, size_t end, v4 *regs,
                                            double *scalar_regs) {
    for (size_t k = 0; k < p->n_consts; ++k) regs[p->param_count + k] = splat(p->consts[k]);
    const size_t provided = p->param_count < n_cols ? p->param_count : n_cols;
    for (size_t c = provided; c < p->param_count; ++c) regs[c] = splat(0.0);
    size_t i = begin;
    for (; i + 4 <= end; i += 4) {
        for (size_t c = 0; c < provided; ++c) {
            /* the reference slices the columns per chunk (batch.rs:66-69): lengths are relative to the chunk start */
            const size_t len = col_lens[c] < end ? col_lens[c] : end;
            if (i + 4 <= len) regs[c] = _mm256_loadu_pd(cols[c] + i);
            else regs[c] = splat(cols[c][len - 1]);
        }
        exec_simd(p, regs);
        for (uint32_t k = 0; k < p->n_results; ++k) _mm256_storeu_pd(outs[k] + i, regs[p->result_regs[k]]);
    }
    /* scalar tail (simd.rs:96-113 -> eval_batch_scalar) */
    if (i < end) {
        if (p->n_consts) memcpy(scalar_regs + p->param_count, p->consts, p->n_consts * sizeof(double));
        for (size_t c = provided; c < p->param_count; ++c) scalar_regs[c] = 0.0;
        for (; i < end; ++i) {
            for (size_t c = 0; c < provided; ++c) {
                const size_t len = col_lens[c];
                scalar_regs[c] = (i < len) ? cols[c][i] : cols[c][len - 1];
            }
            saf_oracle_exec(p, scalar_regs);
            for (uint32_t k = 0; k < p->n_results; ++k) outs[k][i] = scalar_regs[p->result_regs[k]];
        }
    }
}

typedef struct {
    const saf_oracle_program *p;
    const double *const *cols;
    const size_t *col_lens;
    size_t n_cols;
    double *const *outs;
    double *const *state; /* iterate: state columns (cols / outs unused) */
    size_t n_state;
    uint64_t iters;
    size_t n_points;
    atomic_size_t next_chunk;
    size_t n_chunks;
    int failed;
} simd_job;

SAF_SIMD_TARGET static void *simd_worker(void *arg) {
    simd_job *job = (simd_job *)arg;
    const saf_oracle_program *p = job->p;
    const size_t ws = p->workspace_size ? p->workspace_size : 1;
    /* one f64x4 workspace per worker, batch.rs:59-64 */
    v4 *regs = (v4 *)aligned_alloc(32, sizeof(v4) * ws);
    double *sregs = (double *)malloc(sizeof(double) * ws);
    double *tmp = (double *)malloc(sizeof(double) * 4 * (job->n_state ? job->n_state : 1));
    if (!regs || !sregs || !tmp) { job->failed = 1; free(regs); free(sregs); free(tmp); return NULL; }
    for (;;) {
        const size_t c = atomic_fetch_add(&job->next_chunk, 1);
        if (c >= job->n_chunks) break;
        const size_t start = c * SAF_CHUNK_SIZE;
        size_t end = start + SAF_CHUNK_SIZE;
        if (end > job->n_points) end = job->n_points;
        if (!job->state) {
            eval_range_simd(p, job->cols, job->col_lens, job->n_cols, job->outs, start, end, regs, sregs);
            continue;
        }
        /* iterated map: every group of four points runs `iters` rounds in the workspace (points are independent) */
        for (size_t k = 0; k < p->n_consts; ++k) regs[p->param_count + k] = splat(p->consts[k]);
        const size_t provided = p->param_count < job->n_state ? p->param_count : job->n_state;
        for (size_t j = provided; j < p->param_count; ++j) regs[j] = splat(0.0);
        size_t i = start;
        for (; i + 4 <= end; i += 4) {
            v4 st[32];
            for (size_t j = 0; j < job->n_state; ++j) st[j] = _mm256_loadu_pd(job->state[j] + i);
            for (uint64_t it = 0; it < job->iters; ++it) {
                for (size_t j = 0; j < provided; ++j) regs[j] = st[j];
                exec_simd(p, regs);
                for (size_t j = 0; j < job->n_state; ++j) st[j] = regs[p->result_regs[j]];
            }
            for (size_t j = 0; j < job->n_state; ++j) _mm256_storeu_pd(job->state[j] + i, st[j]);
        }
        if (i < end) {
            if (p->n_consts) memcpy(sregs + p->param_count, p->consts, p->n_consts * sizeof(double));
            for (; i < end; ++i) {
                for (size_t j = 0; j < job->n_state; ++j) tmp[j] = job->state[j][i];
                for (uint64_t it = 0; it < job->iters; ++it) {
                    for (size_t j = 0; j < provided; ++j) sregs[j] = tmp[j];
                    for (size_t j = provided; j < p->param_count; ++j) sregs[j] = 0.0;
                    saf_oracle_exec(p, sregs);
                    for (size_t j = 0; j < job->n_state; ++j) tmp[j] = sregs[p->result_regs[j]];
                }
                for (size_t j = 0; j < job->n_state; ++j) job->state[j][i] = tmp[j];
            }
        }
    }
    free(regs);
    free(sregs);
    free(tmp);
    return NULL;
}

static int run_simd_threads(simd_job *job, int n_threads) {
    if (n_threads <= 0) n_threads = saf_oracle_online_cores();
    if ((size_t)n_threads > job->n_chunks) n_threads = (int)job->n_chunks;
    if (n_threads <= 1) {
        simd_worker(job);
        return job->failed ? SAF_ORACLE_ERR_ALLOC : SAF_ORACLE_OK;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    if (!th) return SAF_ORACLE_ERR_ALLOC;
    int started = 0;
    for (int t = 0; t < n_threads; ++t) {
        if (pthread_create(&th[t], NULL, simd_worker, job) != 0) break;
        ++started;
    }
    if (started == 0) simd_worker(job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    return job->failed ? SAF_ORACLE_ERR_ALLOC : SAF_ORACLE_OK;
}

/* batch.rs:37-78 with the f64x4 workspace (what eval_f64 runs for >= 256 points) */
int saf_oracle_run_chunked_simd(const saf_oracle_program *p, const double *const *cols, const size_t *col_lens,
                                size_t n_cols, double *const *outs, size_t n_points, int n_threads) {
    if (!saf_oracle_has_simd()) return saf_oracle_run_chunked(p, cols, col_lens, n_cols, outs, n_points, n_threads);
    if (n_cols == 0) {
        if (n_points == 0) return SAF_ORACLE_OK;
    } else {
        for (size_t c = 0; c < n_cols; ++c)
            if (col_lens[c] != n_points) return SAF_ORACLE_ERR_COLUMN_LENGTH_MISMATCH;
    }
    if (n_points < SAF_CHUNK_SIZE) return saf_oracle_eval_batch(p, cols, col_lens, n_cols, outs, n_points); /* workspace None */
    simd_job job;
    memset(&job, 0, sizeof job);
    job.p = p; job.cols = cols; job.col_lens = col_lens; job.n_cols = n_cols; job.outs = outs;
    job.n_points = n_points;
    atomic_init(&job.next_chunk, 0);
    job.n_chunks = (n_points + SAF_CHUNK_SIZE - 1) / SAF_CHUNK_SIZE;
    return run_simd_threads(&job, n_threads);
}

int saf_oracle_iterate_simd(const saf_oracle_program *p, double *const *state, size_t n_state, size_t n_points,
                            uint64_t iters, int n_threads) {
    if (!saf_oracle_has_simd() || n_state > 32) return saf_oracle_iterate(p, state, n_state, n_points, iters, n_threads);
    if (p->n_results != n_state) return SAF_ORACLE_ERR_INVALID_PROGRAM;
    if (n_points == 0 || iters == 0) return SAF_ORACLE_OK;
    simd_job job;
    memset(&job, 0, sizeof job);
    job.p = p; job.state = state; job.n_state = n_state; job.iters = iters; job.n_points = n_points;
    atomic_init(&job.next_chunk, 0);
    job.n_chunks = (n_points + SAF_CHUNK_SIZE - 1) / SAF_CHUNK_SIZE;
    return run_simd_threads(&job, n_threads);
}

#else /* not x86-64: the scalar path stands in */
int saf_oracle_has_simd(void) { return 0; }
int saf_oracle_run_chunked_simd(const saf_oracle_program *p, const double *const *cols, const size_t *col_lens,
                                size_t n_cols, double *const *outs, size_t n_points, int n_threads) {
    return saf_oracle_run_chunked(p, cols, col_lens, n_cols, outs, n_points, n_threads);
}
int saf_oracle_iterate_simd(const saf_oracle_program *p, double *const *state, size_t n_state, size_t n_points,
                            uint64_t iters, int n_threads) {
    return saf_oracle_iterate(p, state, n_state, n_points, iters, n_threads);
}
#endif
