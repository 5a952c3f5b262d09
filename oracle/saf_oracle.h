/*
 * saf_oracle.h -- CPU restatement of the SymbAnaFis bytecode evaluator.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under symbanafis_b200/ may include, link
 * or call this.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py use it, and only as the checker or as
 * the timed CPU baseline -- never as the product.
 *
 * What it restates (all paths relative to /root/reference):
 *   - the 42-opcode interpreter    src/evaluator/logic/bytecode/execute/engine/macros.rs:1-411
 *   - builtin dispatch             .../execute/engine/builtins.rs:38-139
 *   - sinc / round_to_i32          .../execute/engine/helpers.rs:9-25
 *   - register setup + broadcast   .../execute/engine/scalar.rs:93-108,172-230
 *   - 256-point chunked driver     .../execute/drivers/batch.rs:9,37-78
 *   - flat encoding                .../compile/emit/assemble.rs:9-103
 *   - program validator            .../compile/optimize/helper.rs:98-196
 *   - special functions            src/math/logic/functions/ (all .rs files)
 *
 * Third-party arithmetic that is NOT under /root/reference and is restated
 * from its published algorithm: Rust std f64 methods (glibc libm on Linux for
 * sin/cos/exp/ln/pow/tan/..., compiler-rt __powidf2 for powi, Rust-std
 * formulas for asinh/acosh/atanh), see saf_oracle.c.
 *
 * Parity status: the Rust reference cannot be built in this image (no
 * cargo/rustc), so this oracle is pinned against the reference's own golden
 * vectors / known-answer tests (tests/test_oracle_golden.py lists them with
 * file:line).  Items with no reference test -- fused vs unfused mul_add in the
 * f64x4 body, Builtin1{Sin,Cos,Exp,Ln,Sqrt}, erf for |x|>3, the Rust-std
 * asinh/acosh/atanh formulas -- are "parity unpinned" (SURVEY.md section 8c).
 *
 * Extension over the reference: a program may name several result registers
 * (`result_regs`), evaluated in one pass; the reference has exactly one
 * (`result_reg`, src/evaluator/api.rs:140-141).
 */
#ifndef SAF_ORACLE_H
#define SAF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct saf_oracle_program {
    const uint32_t *flat;       /* flat_bytecode, api.rs:128-129 (trailing End sentinel included) */
    uint64_t n_words;
    const double *consts;       /* constants, api.rs:130-131 */
    uint64_t n_consts;
    const uint32_t *arg_pool;   /* arg_pool, api.rs:132-133 */
    uint64_t n_pool;
    uint32_t param_count;       /* api.rs:138-139 */
    uint32_t workspace_size;    /* api.rs:136-137 */
    const uint32_t *result_regs;/* api.rs:140-141 generalised to k outputs */
    uint32_t n_results;
} saf_oracle_program;

/* status codes (mirrors DiffError variants reachable on the hot path) */
#define SAF_ORACLE_OK 0
#define SAF_ORACLE_ERR_COLUMN_LENGTH_MISMATCH 1 /* DiffError::EvalColumnLengthMismatch, batch.rs:48-50 */
#define SAF_ORACLE_ERR_INVALID_PROGRAM 2        /* validate_program, helper.rs:98-196 */
#define SAF_ORACLE_ERR_ALLOC 3

/* helper.rs:98-196. Writes a message to err (may be NULL). */
int saf_oracle_validate(const saf_oracle_program *p, char *err, size_t errlen);

/* macros.rs:1-357: run the program once on an already set-up register file. */
void saf_oracle_exec(const saf_oracle_program *p, double *regs);

/* scalar.rs:61-108 evaluate(): params shorter than param_count read as 0.
 * Writes n_results values to out. */
int saf_oracle_evaluate(const saf_oracle_program *p, const double *params,
                        size_t n_params, double *out);

/* scalar.rs:172-230 eval_batch_scalar(): per-point loop, a column shorter than
 * n_points is broadcast from its last element, missing columns are 0, extra
 * columns are ignored.  outs[k][i] for k < n_results. */
int saf_oracle_eval_batch(const saf_oracle_program *p, const double *const *cols,
                          const size_t *col_lens, size_t n_cols,
                          double *const *outs, size_t n_points);

/* batch.rs:37-78 run_chunked_evaluator(): every column must have n_points
 * elements; < 256 points runs the scalar loop, otherwise 256-point chunks are
 * spread over n_threads workers (0 = all online cores), one workspace each. */
int saf_oracle_run_chunked(const saf_oracle_program *p, const double *const *cols,
                           const size_t *col_lens, size_t n_cols,
                           double *const *outs, size_t n_points, int n_threads);

/* Iterated map, the loop the example scripts write around eval_f64
 * (examples/python/clifford_benchmark.py:119, aizawa_flow.py:158-165 with the
 * Euler update folded into the program): state[j] <- result[j] for `iters`
 * rounds, n_results must equal n_state.  state columns are updated in place. */
int saf_oracle_iterate(const saf_oracle_program *p, double *const *state,
                       size_t n_state, size_t n_points, uint64_t iters, int n_threads);

/* The reference's f64x4 path (simd.rs:53-114 + the simd arms of macros.rs / builtins.rs:148-299, saf_oracle_simd.c):
 * arithmetic as 4-lane vector operations, transcendentals as four scalar calls, scalar tail, 256-point chunks with one
 * f64x4 workspace per worker.  Same results as the scalar functions above, bit for bit (MulAdd is a true FMA in both).
 * Falls back to the scalar path when the CPU has no AVX2 + FMA (saf_oracle_has_simd() == 0). */
int saf_oracle_has_simd(void);
int saf_oracle_run_chunked_simd(const saf_oracle_program *p, const double *const *cols,
                                const size_t *col_lens, size_t n_cols,
                                double *const *outs, size_t n_points, int n_threads);
int saf_oracle_iterate_simd(const saf_oracle_program *p, double *const *state,
                            size_t n_state, size_t n_points, uint64_t iters, int n_threads);

/* builtins.rs:38-139 exposed for known-answer tests. */
double saf_oracle_builtin1(uint32_t fnop, double x);
double saf_oracle_builtin2(uint32_t fnop, double x1, double x2);
double saf_oracle_builtin3(uint32_t fnop, double x1, double x2, double x3);
double saf_oracle_builtin4(uint32_t fnop, double x1, double x2, double x3, double x4);
/* Rust f64::powi via compiler-rt __powidf2 (macros.rs:374) */
double saf_oracle_powi(double v, int32_t n);

int saf_oracle_online_cores(void);

#ifdef __cplusplus
}
#endif
#endif
