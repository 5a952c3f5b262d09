"""ctypes front end of the CPU oracle (``oracle/saf_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package
``symbanafis_b200`` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsaf_oracle.so")

OK = 0
ERR_COLUMN_LENGTH_MISMATCH = 1
ERR_INVALID_PROGRAM = 2
ERR_ALLOC = 3


class OracleError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


class _Prog(C.Structure):
    _fields_ = [
        ("flat", C.POINTER(C.c_uint32)), ("n_words", C.c_uint64),
        ("consts", C.POINTER(C.c_double)), ("n_consts", C.c_uint64),
        ("arg_pool", C.POINTER(C.c_uint32)), ("n_pool", C.c_uint64),
        ("param_count", C.c_uint32), ("workspace_size", C.c_uint32),
        ("result_regs", C.POINTER(C.c_uint32)), ("n_results", C.c_uint32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``oracle/Makefile``) if the shared object is missing/stale."""
    srcs = [os.path.join(_HERE, f) for f in
            ("saf_oracle.c", "saf_oracle_math.c", "saf_oracle_simd.c", "saf_oracle.h", "saf_oracle_math.h", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libsaf_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        PP = C.POINTER(C.c_void_p)
        L.saf_oracle_validate.argtypes = [C.POINTER(_Prog), C.c_char_p, C.c_size_t]
        L.saf_oracle_validate.restype = C.c_int
        L.saf_oracle_evaluate.argtypes = [C.POINTER(_Prog), C.c_void_p, C.c_size_t, C.c_void_p]
        L.saf_oracle_evaluate.restype = C.c_int
        L.saf_oracle_eval_batch.argtypes = [C.POINTER(_Prog), PP, C.POINTER(C.c_size_t), C.c_size_t, PP, C.c_size_t]
        L.saf_oracle_eval_batch.restype = C.c_int
        L.saf_oracle_run_chunked.argtypes = [C.POINTER(_Prog), PP, C.POINTER(C.c_size_t), C.c_size_t, PP,
                                             C.c_size_t, C.c_int]
        L.saf_oracle_run_chunked.restype = C.c_int
        L.saf_oracle_iterate.argtypes = [C.POINTER(_Prog), PP, C.c_size_t, C.c_size_t, C.c_uint64, C.c_int]
        L.saf_oracle_iterate.restype = C.c_int
        L.saf_oracle_run_chunked_simd.argtypes = L.saf_oracle_run_chunked.argtypes
        L.saf_oracle_run_chunked_simd.restype = C.c_int
        L.saf_oracle_iterate_simd.argtypes = L.saf_oracle_iterate.argtypes
        L.saf_oracle_iterate_simd.restype = C.c_int
        L.saf_oracle_has_simd.restype = C.c_int
        for name, n in (("saf_oracle_builtin1", 1), ("saf_oracle_builtin2", 2),
                        ("saf_oracle_builtin3", 3), ("saf_oracle_builtin4", 4)):
            f = getattr(L, name)
            f.argtypes = [C.c_uint32] + [C.c_double] * n
            f.restype = C.c_double
        L.saf_oracle_powi.argtypes = [C.c_double, C.c_int32]
        L.saf_oracle_powi.restype = C.c_double
        L.saf_oracle_online_cores.restype = C.c_int
        _lib = L
    return _lib


def _ptr_array(arrs: Sequence[np.ndarray]):
    a = (C.c_void_p * max(1, len(arrs)))()
    for i, x in enumerate(arrs):
        a[i] = x.ctypes.data
    return a


class OracleProgram:
    """Holds the numpy buffers of a program alive and exposes the oracle entry points."""

    def __init__(self, flat, consts, arg_pool, param_count: int, workspace_size: int,
                 result_regs: Sequence[int], validate: bool = True):
        self.flat = np.ascontiguousarray(flat, dtype=np.uint32)
        self.consts = np.ascontiguousarray(consts, dtype=np.float64)
        self.arg_pool = np.ascontiguousarray(arg_pool, dtype=np.uint32)
        self.result_regs = np.ascontiguousarray(result_regs, dtype=np.uint32)
        self.param_count = int(param_count)
        self.workspace_size = int(workspace_size)
        self.n_results = len(self.result_regs)
        self._c = _Prog(
            self.flat.ctypes.data_as(C.POINTER(C.c_uint32)), self.flat.size,
            self.consts.ctypes.data_as(C.POINTER(C.c_double)), self.consts.size,
            self.arg_pool.ctypes.data_as(C.POINTER(C.c_uint32)), self.arg_pool.size,
            self.param_count, self.workspace_size,
            self.result_regs.ctypes.data_as(C.POINTER(C.c_uint32)), self.n_results)
        if validate:
            self.validate()

    @classmethod
    def from_program(cls, prog, validate: bool = True) -> "OracleProgram":
        return cls(prog.flat, prog.consts, prog.arg_pool, prog.param_count, prog.workspace_size,
                   prog.result_regs, validate=validate)

    def validate(self) -> None:
        buf = C.create_string_buffer(256)
        rc = lib().saf_oracle_validate(C.byref(self._c), buf, 256)
        if rc != OK:
            raise OracleError(rc, buf.value.decode())

    def evaluate(self, params: Sequence[float]) -> np.ndarray:
        p = np.ascontiguousarray(params, dtype=np.float64)
        out = np.empty(self.n_results, dtype=np.float64)
        rc = lib().saf_oracle_evaluate(C.byref(self._c), p.ctypes.data, p.size, out.ctypes.data)
        if rc != OK:
            raise OracleError(rc)
        return out

    def _run(self, fn, cols: Sequence[np.ndarray], n_points: int, *extra) -> List[np.ndarray]:
        cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
        lens = (C.c_size_t * max(1, len(cols)))(*[c.size for c in cols])
        outs = [np.empty(n_points, dtype=np.float64) for _ in range(self.n_results)]
        rc = fn(C.byref(self._c), _ptr_array(cols), lens, len(cols), _ptr_array(outs), n_points, *extra)
        if rc != OK:
            raise OracleError(rc, "column length mismatch" if rc == ERR_COLUMN_LENGTH_MISMATCH else "")
        return outs

    def eval_batch(self, cols: Sequence[np.ndarray], n_points: Optional[int] = None) -> List[np.ndarray]:
        """``CompiledEvaluator::eval_batch(.., None)``: scalar loop with last-value broadcast."""
        if n_points is None:
            n_points = len(cols[0]) if len(cols) else 1
        return self._run(lib().saf_oracle_eval_batch, cols, n_points)

    def run_chunked(self, cols: Sequence[np.ndarray], n_points: Optional[int] = None,
                    n_threads: int = 0) -> List[np.ndarray]:
        """``run_chunked_evaluator``: equal-length columns, 256-point chunks over threads."""
        if n_points is None:
            n_points = len(cols[0]) if len(cols) else 1
        return self._run(lib().saf_oracle_run_chunked, cols, n_points, n_threads)

    def run_chunked_simd(self, cols: Sequence[np.ndarray], n_points: Optional[int] = None,
                         n_threads: int = 0) -> List[np.ndarray]:
        """``run_chunked_evaluator`` with the f64x4 workspace (simd.rs:53-114): what ``eval_f64`` runs."""
        if n_points is None:
            n_points = len(cols[0]) if len(cols) else 1
        return self._run(lib().saf_oracle_run_chunked_simd, cols, n_points, n_threads)

    def iterate(self, state: Sequence[np.ndarray], iters: int, n_threads: int = 0, simd: bool = False,
                in_place: bool = False) -> List[np.ndarray]:
        st = list(state) if in_place else [np.array(s, dtype=np.float64, copy=True) for s in state]
        n = st[0].size if st else 0
        fn = lib().saf_oracle_iterate_simd if simd else lib().saf_oracle_iterate
        rc = fn(C.byref(self._c), _ptr_array(st), len(st), n, iters, n_threads)
        if rc != OK:
            raise OracleError(rc)
        return st


def builtin(fnop: int, *args: float) -> float:
    return getattr(lib(), f"saf_oracle_builtin{len(args)}")(fnop, *args)


def powi(v: float, n: int) -> float:
    return lib().saf_oracle_powi(v, n)


def online_cores() -> int:
    return lib().saf_oracle_online_cores()


def has_simd() -> bool:
    """AVX2 + FMA available: the f64x4 body runs as vectors (else it falls back to the scalar path)."""
    return bool(lib().saf_oracle_has_simd())
