// Rust side of the C ABI in include/saf_b200.h.  NOT compiled in this repository's image
// (no cargo / rustc): shipped as the source a SymbAnaFis maintainer drops into
// src/evaluator/logic/bytecode/execute/drivers/ behind the cargo feature `b200`.
use crate::core::DiffError;
use crate::evaluator::CompiledEvaluator;
use std::os::raw::{c_char, c_int, c_uint, c_void};

#[repr(C)] pub struct SafProgram { _private: [u8; 0] }

extern "C" {
    fn saf_program_create(flat: *const u32, n_words: usize, consts: *const f64, n_consts: usize,
                          pool: *const u32, n_pool: usize, param_count: u32, workspace_size: u32,
                          result_regs: *const u32, n_results: u32, out: *mut *mut SafProgram) -> c_int;
    fn saf_program_destroy(p: *mut SafProgram);
    fn saf_program_specialize(p: *mut SafProgram) -> c_int;
    fn saf_program_intern(flat: *const u32, n_words: usize, consts: *const f64, n_consts: usize,
                          pool: *const u32, n_pool: usize, param_count: u32, workspace_size: u32,
                          result_regs: *const u32, n_results: u32, out: *mut *mut SafProgram) -> c_int;
    fn saf_eval_host(p: *const SafProgram, cols: *const *const f64, col_lens: *const usize, n_cols: usize,
                     outs: *const *mut f64, n_points: usize, flags: c_uint) -> c_int;
    fn saf_last_error() -> *const c_char;
}
const SAF_ERR_COLUMN_LENGTH_MISMATCH: c_int = 1;

/// Device twin of a `CompiledEvaluator`; created once, reused, `Send + Sync` like the evaluator itself.
pub struct B200Program(*mut SafProgram);
unsafe impl Send for B200Program {}
unsafe impl Sync for B200Program {}
impl Drop for B200Program { fn drop(&mut self) { unsafe { saf_program_destroy(self.0) } } }

impl B200Program {
    pub fn new(ev: &CompiledEvaluator) -> Result<Self, DiffError> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe {
            saf_program_create(ev.flat_bytecode.as_ptr(), ev.flat_bytecode.len(),
                               ev.constants.as_ptr(), ev.constants.len(),
                               ev.arg_pool.as_ptr(), ev.arg_pool.len(),
                               ev.param_count as u32, ev.workspace_size as u32,
                               &ev.result_reg, 1, &mut h)
        };
        if rc != 0 { return Err(last_error(rc)); }
        Ok(Self(h))
    }

    /// Compiles the program into a straight-line sm_100a kernel of its own (NVRTC, once; tenths of a second to a few
    /// seconds): every later evaluation through this handle launches it instead of the interpreter and returns the
    /// same bits.  An `Err` (no libnvrtc, program above 20 000 device instructions) leaves the handle on the
    /// interpreter, fully usable.
    pub fn specialize(&self) -> Result<(), DiffError> {
        let rc = unsafe { saf_program_specialize(self.0) };
        if rc != 0 { return Err(last_error(rc)); }
        Ok(())
    }
}

/// `eval_f64` compiles its expressions inside every call (drivers/batch.rs:28-30): there is no long-lived
/// `CompiledEvaluator` to attach a `B200Program` to.  That path takes its handle from the library's content-addressed
/// cache instead (`saf_program_intern`: same payload, same handle; owned by the cache, never destroyed here).
pub fn interned_program(ev: &CompiledEvaluator) -> Result<*const SafProgram, DiffError> {
    let mut h = std::ptr::null_mut();
    let rc = unsafe {
        saf_program_intern(ev.flat_bytecode.as_ptr(), ev.flat_bytecode.len(),
                           ev.constants.as_ptr(), ev.constants.len(),
                           ev.arg_pool.as_ptr(), ev.arg_pool.len(),
                           ev.param_count as u32, ev.workspace_size as u32,
                           &ev.result_reg, 1, &mut h)
    };
    if rc != 0 { return Err(last_error(rc)); }
    Ok(h as *const SafProgram)
}

/// Same signature and error behaviour as `run_chunked_evaluator` (batch.rs:37-78).
pub fn run_chunked_evaluator_b200(prog: &B200Program, columns: &[&[f64]], output: &mut [f64])
    -> Result<(), DiffError> {
    let ptrs: Vec<*const f64> = columns.iter().map(|c| c.as_ptr()).collect();
    let lens: Vec<usize> = columns.iter().map(|c| c.len()).collect();
    let out = [output.as_mut_ptr()];
    let rc = unsafe { saf_eval_host(prog.0, ptrs.as_ptr(), lens.as_ptr(), columns.len(),
                                    out.as_ptr(), output.len(), 0) };
    match rc { 0 => Ok(()), SAF_ERR_COLUMN_LENGTH_MISMATCH => Err(DiffError::EvalColumnLengthMismatch),
              _ => Err(last_error(rc)) }
}

fn last_error(_rc: c_int) -> DiffError {
    let msg = unsafe { std::ffi::CStr::from_ptr(saf_last_error()) }.to_string_lossy().into_owned();
    DiffError::invalid_syntax(msg) // or a new DiffError::Backend(String) variant
}
