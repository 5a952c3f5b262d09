/*
 * saf_b200.h -- C ABI of the B200-native batch evaluator for SymbAnaFis compiled expressions.
 *
 * This is the drop-in boundary for ONE hot path of the reference (paths relative to the
 * SymbAnaFis repository):
 *
 *   run_chunked_evaluator(&CompiledEvaluator, &[&[f64]], &mut [f64]) -> Result<(), DiffError>
 *       src/evaluator/logic/bytecode/execute/drivers/batch.rs:37-78
 *   CompiledEvaluator::eval_batch(&self, &[&[f64]], &mut [f64], Option<&mut [f64x4]>)
 *       src/evaluator/logic/bytecode/execute/engine/scalar.rs:150-169
 *
 * reached from eval_f64 (src/evaluator/api.rs:352-376 -> batch.rs:33), from the numeric fast path of
 * evaluate_parallel (.../drivers/parallel.rs:471) and from the PyO3 wrappers
 * (src/bindings/python/evaluator.rs:73-113, parallel.rs:177-256).
 *
 * The library takes the reference's own program payload -- the fields of `CompiledEvaluator`
 * (src/evaluator/api.rs:125-142): the flat u32 bytecode produced by assemble_flat_bytecode
 * (.../compile/emit/assemble.rs:9-103, 42 opcodes of instruction.rs:295-395, FnOp ids of
 * functions.rs:38-117), the f64 constant pool, the u32 arg pool, param_count, workspace_size and
 * result_reg -- lowers it once to a device program and runs it in a hand-written sm_100a
 * interpreter kernel.  There is no CPU fallback: without a CUDA device every evaluation entry
 * point returns SAF_ERR_CUDA.
 *
 * Plain C types only; no exceptions or panics cross this boundary; every function returns an
 * int status (0 = ok) and saf_last_error() gives the thread-local message of the last failure.
 * A program handle is immutable after creation and may be used from many host threads at once
 * (the reference evaluator is Send + Sync, api.rs:113-124).
 */
#ifndef SAF_B200_H
#define SAF_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SAF_API __attribute__((visibility("default")))
#else
#define SAF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------ */
#define SAF_OK 0
/* DiffError::EvalColumnLengthMismatch (batch.rs:48-50, message core/helpers/logic/error.rs:363-365) */
#define SAF_ERR_COLUMN_LENGTH_MISMATCH 1
/* the always-on equivalent of validate_program (compile/optimize/helper.rs:98-196) rejected the program */
#define SAF_ERR_INVALID_PROGRAM 2
#define SAF_ERR_ALLOC 3
#define SAF_ERR_CUDA 4             /* CUDA runtime failure or no device; message has the CUDA string */
#define SAF_ERR_INVALID_ARGUMENT 5 /* NULL pointers, zero-length broadcast column, ... */
#define SAF_ERR_UNSUPPORTED 6      /* program exceeds a device-program limit */

/* ---- evaluation flags --------------------------------------------------------------------- */
/* Default = run_chunked_evaluator semantics: every supplied column must hold exactly n_points
 * values (batch.rs:48-50).  With SAF_EVAL_BROADCAST the call has CompiledEvaluator::eval_batch
 * semantics instead: a column shorter than n_points is read as its last element from there on
 * (scalar.rs:203-207).  In both modes a missing parameter column reads as 0.0 and extra columns
 * are ignored (scalar.rs:182-197). */
#define SAF_EVAL_BROADCAST 1u

typedef struct saf_program saf_program; /* opaque */

typedef struct saf_program_info {
    uint32_t param_count;        /* inputs per point */
    uint32_t n_results;          /* outputs per point */
    uint32_t n_ref_instructions; /* instructions in the reference bytecode (all fused programs) */
    uint32_t n_dev_instructions; /* device instructions after value numbering / dead-code removal */
    uint32_t n_slots;            /* live-value slots per point in the on-chip register file */
    uint32_t n_consts;           /* device constant table entries */
    uint32_t n_acc_forwards;     /* operands served from the accumulator instead of a slot */
    uint32_t quirk_flags;        /* SAF_QUIRK_* */
    double fp64_slots;           /* algorithmic FP64 issue slots per point-eval (SURVEY.md 8d cost table) */
    double bytes_per_point;      /* algorithmic HBM bytes per point-eval = 8*(cols read + cols written) */
} saf_program_info;

/* The program contains Builtin1 with FnOp Sin/Cos/Exp/Ln/Sqrt, which the reference's release build
 * evaluates to NaN (engine/builtins.rs:21-26,84); this library does the same. */
#define SAF_QUIRK_BUILTIN1_UNREACHABLE 1u

/* ---- program lifetime --------------------------------------------------------------------- */

/* Replaces: holding a `CompiledEvaluator` (api.rs:125-142).  Copies everything it needs; the
 * caller's buffers may be freed after the call.  `result_regs` generalises `result_reg` to k
 * outputs evaluated in one pass (k = 1 for a reference program). */
SAF_API int saf_program_create(const uint32_t *flat_bytecode, size_t n_words,
                       const double *constants, size_t n_constants,
                       const uint32_t *arg_pool, size_t n_arg_pool,
                       uint32_t param_count, uint32_t workspace_size,
                       const uint32_t *result_regs, uint32_t n_results,
                       saf_program **out);

/* Fuses k programs compiled independently over the SAME parameter list (what eval_f64 receives
 * when a caller passes [dx, dy, dz] with identical columns, api.rs:365-375) into one multi-output
 * program; common subexpressions across the k instruction streams are evaluated once.  Output j
 * of the fused program is output 0.. of program j in order. */
SAF_API int saf_program_fuse(const saf_program *const *programs, size_t n_programs, saf_program **out);

/* Expression strings -> program, without Python or Rust (csrc/compile.cpp): the k
 * expressions are parsed (the reference's grammar subset: + - * / ^ **, unary minus, calls, implicit multiplication;
 * src/parser/logic/pratt.rs:87-120), lowered with the reference's instruction-selection rules
 * (src/evaluator/logic/bytecode/compile/codegen/lower/ (all files): MulAdd / MulSub / NegMulAdd, Square .. InvPow3_2 / Powi,
 * ExpSqr / ExpSqrNeg / Builtin1{ExpNeg}, RecipExpm1, Sinc, u/c -> u*(1/c), Sin + Cos -> SinCos), value-numbered across
 * the k outputs, register-allocated (emit/reg_alloc.rs) and encoded (emit/assemble.rs) into ONE multi-output program
 * over `params` -- what CompiledEvaluator::compile produces, one register file for all outputs.  Compile errors
 * (UnboundVariable, UnsupportedFunction: api.rs:402-408) -> SAF_ERR_INVALID_PROGRAM, parse errors ->
 * SAF_ERR_INVALID_ARGUMENT, message in saf_last_error(). */
SAF_API int saf_compile_exprs(const char *const *exprs, size_t n_exprs, const char *const *params, size_t n_params,
                              saf_program **out);

/* The reference-format payload (api.rs:125-142 fields) a handle was created from: component `index` of a fused handle,
 * 0 otherwise.  Any pointer may be NULL; the n_* outputs always receive the full sizes. */
SAF_API int saf_program_reference_payload(const saf_program *prog, size_t index, uint32_t *flat, size_t cap_words,
                                          size_t *n_words, double *consts, size_t cap_consts, size_t *n_consts,
                                          uint32_t *arg_pool, size_t cap_pool, size_t *n_pool, uint32_t *param_count,
                                          uint32_t *workspace_size, uint32_t *result_regs, size_t cap_results,
                                          size_t *n_results);

SAF_API void saf_program_destroy(saf_program *prog);

/* Same arguments as saf_program_create, but the handle comes from a per-process cache keyed by the CONTENT of the
 * payload: the reference compiles inside every eval_f64 call (drivers/batch.rs:28-30), so a caller that rebuilds its
 * CompiledEvaluator per call still reuses the lowered program, its packed images and their device copies.  The handle
 * belongs to the cache: do not pass it to saf_program_destroy; it stays valid until saf_program_cache_clear (which
 * the caller must not run concurrently with evaluations on interned handles).  Thread-safe.  The cache holds at most
 * 4096 programs; beyond that the call fails with SAF_ERR_ALLOC. */
SAF_API int saf_program_intern(const uint32_t *flat_bytecode, size_t n_words,
                       const double *constants, size_t n_constants,
                       const uint32_t *arg_pool, size_t n_arg_pool,
                       uint32_t param_count, uint32_t workspace_size,
                       const uint32_t *result_regs, uint32_t n_results,
                       saf_program **out);
SAF_API void saf_program_cache_clear(void);

SAF_API int saf_program_get_info(const saf_program *prog, saf_program_info *info);

/* ---- specialised kernels -------------------------------------------------------------------
 * The interpreter runs any program the moment it is created.  A caller that will evaluate ONE program over many
 * points -- the case of CompiledEvaluator (compile once, api.rs:125-142, evaluate many times) -- can additionally have
 * the program written out as a straight-line sm_100a kernel of its own: every live value a register, every constant an
 * immediate, no instruction fetch, no dispatch (symbanafis_b200/csrc/spec.cpp generates the text from the same device
 * program the interpreter runs, NVRTC compiles it; the statements are the interpreter's handlers, so both machines
 * return the same bits).  After a successful call every saf_eval_* / saf_iterate_* of this program launches the
 * specialised kernel.  Costs a run-time compilation (tens of milliseconds to seconds, saf_specialized_info.compile_ms);
 * needs libnvrtc >= 12.8 at run time (SAF_ERR_UNSUPPORTED with the reason otherwise -- the program stays usable).
 * Programs above 20000 device instructions are not specialised (SAF_ERR_UNSUPPORTED).
 * Environment: SAF_SPECIALIZE=0 keeps every launch on the interpreter, =1 specialises every program at first launch,
 * =auto specialises a program at the first launch after it has run 2e10 instruction-points on the interpreter (50-100 ms
 * of work: the order of one compilation) -- for callers that cannot say in advance which programs will be reused.
 * SAF_SPEC_CACHE_DIR=<directory> keeps compiled kernels on disk (keyed by content): the next process skips NVRTC. */
SAF_API int saf_program_specialize(saf_program *prog);

typedef struct saf_specialized_info {
    int compiled;            /* 1 after a successful saf_program_specialize */
    int can_iterate;         /* saf_iterate_* uses the specialised kernel too (results pair with parameters) */
    int points_per_thread;   /* 1, 2 or 4 */
    int block_threads;
    int min_blocks;          /* __launch_bounds__ resident-block target the register allocation was given */
    int registers;           /* per thread; -1 until the kernel has been loaded on the current device */
    int local_bytes;         /* spill / local memory per thread; -1 until loaded */
    int resident_blocks;     /* per SM on the current device; -1 until loaded */
    double generate_ms;      /* writing the kernel text */
    double compile_ms;       /* NVRTC */
    size_t cubin_bytes;
    size_t source_bytes;
    /* the SMALL-BATCH shape: one point per thread in 128-thread blocks, compiled the first time a launch brings fewer
     * points than 12 warps per SM at the shape above (50 k double pendulums: 2.9 ms against 4.3 ms) */
    int small_compiled;
    int small_points_per_thread;
    int small_block_threads;
    int small_registers;     /* -1 until loaded on the current device */
    double small_compile_ms;
    /* short programs that read few columns carry a second entry point (saf_spec_pf) that single passes (saf_eval_*)
     * launch: it loads the columns of a block's next tile before it evaluates the current one */
    int has_prefetch_entry;
    int prefetch_registers;  /* -1 until loaded / when absent */
} saf_specialized_info;
SAF_API int saf_program_specialized_info(const saf_program *prog, saf_specialized_info *info);
/* The generated CUDA C++ text (for inspection) of the large-batch (small_batch = 0) or the small-batch shape.
 * Returns the bytes needed (excluding NUL); writes at most cap. */
SAF_API size_t saf_program_specialized_source(const saf_program *prog, int small_batch, char *buf, size_t cap);
/* 1 when a usable NVRTC can be loaded in this process, else 0 with the reason in saf_last_error(). */
SAF_API int saf_specialize_available(void);

/* Human-readable listing of the lowered device program (cf. CompiledEvaluator::disassemble,
 * api.rs:190-267).  Returns the number of bytes needed (excluding NUL); writes at most cap. */
SAF_API size_t saf_program_disassemble(const saf_program *prog, char *buf, size_t cap);

/* Copies out the lowered device program (64-bit words incl. the end marker, constant table,
 * slot of each parameter [0xffff = unused], S/K field of each result) for tooling and for the CPU
 * emulator the test-suite uses to check the lowering without a GPU.  Any pointer may be NULL;
 * n_code / n_consts always receive the full sizes. */
SAF_API int saf_program_export(const saf_program *prog, uint64_t *code, size_t cap_code, size_t *n_code,
                       double *consts, size_t cap_consts, size_t *n_consts,
                       uint16_t *param_slot, uint32_t *result_field);

/* Copies out the PACKED micro-instruction image the kernel executes for `points_per_thread` in {1, 2, 4, 8}
 * (layout: symbanafis_b200/csrc/uop.h): constants first, micro-instructions from byte `code_off`.
 * param_off / result_off are byte offsets (-1 = parameter never read), result_is_k[k] = 1 when result k is
 * a constant whose offset points into the image.  Iterated maps are double buffered when possible
 * (*double_buffered = 1): the image then holds the program twice, the second copy with parameter and result slots
 * swapped, each U_END continuing with the other copy, and after an EVEN number of iterations the results sit at
 * result_off_even.  points_per_thread = 104 exports the image of the opt-in REGISTER machine instead (four points per
 * thread, live values in real registers; layout: symbanafis_b200/csrc/ruop.h; never double buffered -- the image
 * loops in-kernel by itself when results pair with parameters); points_per_thread = 1000 * P + T (P = 2 or 4, T a multiple of 32,
 * 160..512) the P-points-per-thread image for one wide block of T threads per SM.  For the CPU emulators of the test-suite and for
 * tooling.  Any pointer may be NULL; n_words always receives the full size. */
SAF_API int saf_program_export_packed(const saf_program *prog, int points_per_thread,
                              uint64_t *image, size_t cap_words, size_t *n_words,
                              uint32_t *code_off, uint32_t *n_slots, int32_t *x_off /* [3]: x0, x1, acc */,
                              int32_t *param_off, int32_t *result_off, uint8_t *result_is_k,
                              int32_t *result_off_even, int *double_buffered);

/* ---- evaluation --------------------------------------------------------------------------- */

/* Replaces run_chunked_evaluator / eval_batch with DEVICE-resident struct-of-arrays columns.
 * cols[j] / outs[k] are device pointers on the current CUDA device; col_lens[j] in elements.
 * Asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream).  n_points == 0 is Ok. */
SAF_API int saf_eval_device(const saf_program *prog,
                    const double *const *cols, const size_t *col_lens, size_t n_cols,
                    double *const *outs, size_t n_points,
                    unsigned flags, void *cuda_stream);

/* Same call with HOST buffers (exactly what the Rust/Python callers hold): chunks the points,
 * overlaps host->device copies, the kernel and device->host copies on internal streams of the
 * current device, and returns when outs[] is complete. */
SAF_API int saf_eval_host(const saf_program *prog,
                  const double *const *cols, const size_t *col_lens, size_t n_cols,
                  double *const *outs, size_t n_points, unsigned flags);

/* Host buffers, points sharded contiguously over `n_devices` GPUs of this process (device ids in
 * `devices`, NULL = 0..n_devices-1), one stream set per device, no collective. */
SAF_API int saf_eval_host_multi(const saf_program *prog,
                        const double *const *cols, const size_t *col_lens, size_t n_cols,
                        double *const *outs, size_t n_points, unsigned flags,
                        const int *devices, int n_devices);

/* Iterated map: state[j] <- result j of the program applied to state, `iters` times, entirely
 * on-chip (no HBM traffic between iterations).  Replaces the Python loops the reference's
 * examples write around eval_f64 (examples/python/clifford_benchmark.py:119,
 * aizawa_flow.py:158-165, double_pendulum_benchmark.py:317-365).  Requires
 * n_results == n_state == param_count.  State columns are updated in place. */
SAF_API int saf_iterate_device(const saf_program *prog, double *const *state, size_t n_state,
                       size_t n_points, uint64_t iters, void *cuda_stream);
SAF_API int saf_iterate_host(const saf_program *prog, double *const *state, size_t n_state,
                     size_t n_points, uint64_t iters);

/* ---- misc --------------------------------------------------------------------------------- */
SAF_API const char *saf_last_error(void); /* thread-local, never NULL */
SAF_API int saf_device_count(int *count);
/* Measures the FP64 pipe peak of the current device (independent DFMA chains on every SM) in
 * lane-instructions per second (x2 = FLOP/s).  This is the compute roofline denominator of this path;
 * it is not part of the evaluation path. */
SAF_API int saf_measure_fp64_peak(double *dfma_lanes_per_second);
/* Times the Clifford map hard-coded around the same lean sin/cos kernels (no interpreter, state in registers,
 * same launch shape): the "ideal kernel" reference point bench.py prints next to the interpreter's number.
 * Not part of the evaluation path. */
SAF_API int saf_measure_clifford_probe(size_t n_points, int iters, double *seconds);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
SAF_API uint64_t saf_kernel_launch_count(void);
SAF_API const char *saf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SAF_B200_H */
