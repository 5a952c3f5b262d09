"""Python surface of the evaluator, mirroring the reference's PyO3 API for the hot path.

Same names, argument meaning, return conventions and error types as

* ``CompiledEvaluator(expr, params=None, context=None)`` with ``evaluate``, ``eval_batch``,
  ``param_names``, ``param_count``, ``instruction_count``, ``workspace_size``, ``disassemble``
  -- ``src/bindings/python/evaluator.rs:15-139``
* ``eval_f64(expressions, var_names, data)`` -- ``src/bindings/python/parallel.rs:177-256``
  (Rust side ``src/evaluator/api.rs:352-376``)
* ``evaluate_parallel(expressions, variables, values)``, numeric fast path only --
  ``src/evaluator/logic/bytecode/execute/drivers/parallel.rs:436-487``

but every evaluation runs in the sm_100a interpreter kernel behind the C ABI
(``include/saf_b200.h``).  There is no CPU path here.

Error mapping (``src/bindings/python/error.rs:14-44``): column/dimension errors -> ``ValueError``,
compile errors -> ``RuntimeError``, wrong input type -> ``TypeError``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from .compiler import CompileError, compile_exprs
from .expr import KNOWN_CONSTANTS, Expr, parse
from .isa import Program

ArrayLike = Union[np.ndarray, Sequence[float]]


def _extract(col, what: str = "Input") -> Tuple[np.ndarray, bool]:
    """``extract_data_input`` + ``as_slice`` (evaluator.rs:146-191): (f64 vector, was_ndarray)."""
    if isinstance(col, np.ndarray):
        if col.ndim != 1 or col.dtype != np.float64:
            raise TypeError("Input must be a NumPy array or list of floats")
        if not col.flags.c_contiguous:
            raise ValueError("All input arrays must be C-contiguous and f64: the array is not contiguous")
        return col, True
    if isinstance(col, (list, tuple)):
        try:
            return np.asarray([float(v) for v in col], dtype=np.float64), False
        except (TypeError, ValueError):
            raise TypeError("Input must be a NumPy array or list of floats") from None
    raise TypeError("Input must be a NumPy array or list of floats")


def _as_expr(e) -> Expr:
    if isinstance(e, Expr):
        return e
    if isinstance(e, str):
        return parse(e)
    raise TypeError("expression must be an Expr or a string")


def _raise_from_saf(err: _lib.SafError):
    if err.code == _lib.SAF_ERR_COLUMN_LENGTH_MISMATCH:
        # DiffError::EvalColumnLengthMismatch -> PyValueError (error.rs)
        raise ValueError("All evaluation columns must have the same length") from None
    if err.code in (_lib.SAF_ERR_INVALID_PROGRAM, _lib.SAF_ERR_UNSUPPORTED):
        raise RuntimeError(err.message) from None
    if err.code == _lib.SAF_ERR_INVALID_ARGUMENT:
        raise ValueError(err.message) from None
    raise err


class CompiledEvaluator:
    """Compiled expression (possibly several outputs) bound to a device program handle."""

    def __init__(self, expr, params: Optional[Sequence[str]] = None, context=None):
        if context is not None:
            raise RuntimeError("UnsupportedExpression: user-function contexts are not supported")
        exprs = [_as_expr(e) for e in expr] if isinstance(expr, (list, tuple)) else [_as_expr(expr)]
        if params is None:  # compile_auto: alphabetical, known constants excluded (api.rs:461-473)
            names = sorted({v for e in exprs for v in e.variables()} - set(KNOWN_CONSTANTS))
        else:
            names = [str(p) for p in params]
        try:
            prog = compile_exprs(exprs, names)
        except CompileError as ce:
            raise RuntimeError(str(ce)) from None
        self._init_from_program(prog)

    @classmethod
    def from_program(cls, prog: Program) -> "CompiledEvaluator":
        """Wrap an existing program payload (e.g. bytecode dumped from the Rust reference)."""
        self = cls.__new__(cls)
        self._init_from_program(prog)
        return self

    def _init_from_program(self, prog: Program) -> None:
        self.program = prog
        try:
            self._h = _lib.DeviceProgramHandle.create(prog.flat, prog.consts, prog.arg_pool, prog.param_count,
                                                      prog.workspace_size, prog.result_regs)
        except _lib.SafError as e:
            _raise_from_saf(e)
        self.n_results = len(prog.result_regs)

    # -- introspection (evaluator.rs:115-138) -------------------------------------------------
    def param_names(self) -> List[str]:
        return list(self.program.param_names)

    def param_count(self) -> int:
        return self.program.param_count

    def instruction_count(self) -> int:
        return self.program.instruction_count()

    def workspace_size(self) -> int:
        return self.program.workspace_size

    def disassemble(self) -> str:
        return self.program.disassemble()

    def device_disassemble(self) -> str:
        return self._h.disassemble()

    def device_info(self):
        return self._h.info()

    @property
    def handle(self) -> _lib.DeviceProgramHandle:
        return self._h

    # -- evaluation -------------------------------------------------------------------------------
    def evaluate(self, input) -> float:  # noqa: A002 - the reference's name
        """Single point (evaluator.rs:57-61); missing parameters read as 0 (scalar.rs:93-108)."""
        data, _ = _extract(input)
        cols = [np.array([v]) for v in data[: self.program.param_count]]
        outs = self._run_host(cols, 1, broadcast=True)
        return float(outs[0][0])

    def eval_batch(self, columns: Sequence[ArrayLike]):
        """``columns[var][point]`` -> results; ndarray out iff the FIRST column is an ndarray
        (evaluator.rs:73-113).  ``eval_batch`` semantics: short columns broadcast their last value."""
        extracted = [_extract(c) for c in columns]
        cols = [c for c, _ in extracted]
        n_points = 1 if not cols else cols[0].size
        use_numpy = bool(extracted) and extracted[0][1]
        if n_points == 0:
            outs = [np.empty(0) for _ in range(self.n_results)]
        else:
            outs = self._run_host(cols, n_points, broadcast=True)
        if self.n_results == 1:
            return outs[0] if use_numpy else outs[0].tolist()
        return np.stack(outs) if use_numpy else [o.tolist() for o in outs]

    def _run_host(self, cols: Sequence[np.ndarray], n_points: int, broadcast: bool,
                  devices: Optional[Sequence[int]] = None) -> List[np.ndarray]:
        outs = [np.empty(n_points, dtype=np.float64) for _ in range(self.n_results)]
        try:
            self._h.eval_host(cols, outs, n_points, broadcast=broadcast, devices=devices)
        except _lib.SafError as e:
            _raise_from_saf(e)
        return outs

    def run_chunked(self, columns: Sequence[np.ndarray], n_points: Optional[int] = None,
                    devices: Optional[Sequence[int]] = None) -> List[np.ndarray]:
        """``run_chunked_evaluator`` (batch.rs:37-78): strict equal-length columns, host buffers."""
        cols = [np.ascontiguousarray(c, dtype=np.float64) for c in columns]
        if n_points is None:
            n_points = cols[0].size if cols else 1
        return self._run_host(cols, n_points, broadcast=False, devices=devices)

    def iterate(self, state: Sequence[np.ndarray], iters: int) -> List[np.ndarray]:
        """state_j <- result_j, ``iters`` times on chip; returns new host arrays."""
        st = [np.array(s, dtype=np.float64, copy=True) for s in state]
        if len(st) != self.program.param_count or len(st) != self.n_results:
            raise ValueError("iterate needs one state column per parameter and per result")
        # the C ABI takes ONE length for all state pointers: a shorter later array would be read and written out of bounds
        if any(a.ndim != 1 for a in st) or any(a.size != st[0].size for a in st):
            raise ValueError("All evaluation columns must have the same length")
        try:
            self._h.iterate_host(st, iters)
        except _lib.SafError as e:
            _raise_from_saf(e)
        return st

    # -- device-resident columns (torch tensors) --------------------------------------------------------
    def eval_device(self, columns, outs=None, stream=None):
        """Columns are 1-D float64 CUDA tensors; returns (or fills) one CUDA tensor per result."""
        import torch
        n = columns[0].numel() if columns else 1
        dev = columns[0].device if columns else torch.device("cuda", torch.cuda.current_device())
        _check_device_columns(columns, dev, None, "device columns")
        if any(c.numel() != n for c in columns):
            raise ValueError("All evaluation columns must have the same length")
        if outs is None:
            outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(self.n_results)]
        else:  # user buffers are written by the kernel: a short or foreign one is an out-of-bounds device write
            if len(outs) != self.n_results:
                raise ValueError(f"expected {self.n_results} output tensors, got {len(outs)}")
            _check_device_columns(outs, dev, n, "output tensors")
        s = (stream or torch.cuda.current_stream(dev)).cuda_stream
        try:
            with torch.cuda.device(dev):
                self._h.eval_device([c.data_ptr() for c in columns], [c.numel() for c in columns],
                                    [o.data_ptr() for o in outs], n, broadcast=False, stream=s)
        except _lib.SafError as e:
            _raise_from_saf(e)
        return outs

    def iterate_device(self, state, iters: int, stream=None):
        """In-place iterated map on CUDA tensors."""
        import torch
        if len(state) != self.program.param_count or len(state) != self.n_results:
            raise ValueError("iterate needs one state tensor per parameter and per result")
        n = state[0].numel()
        dev = state[0].device
        _check_device_columns(state, dev, n, "state tensors")
        s = (stream or torch.cuda.current_stream(dev)).cuda_stream
        try:
            with torch.cuda.device(dev):
                self._h.iterate_device([t.data_ptr() for t in state], n, iters, stream=s)
        except _lib.SafError as e:
            _raise_from_saf(e)
        return state


def _check_device_columns(tensors, dev, n, what: str) -> None:
    """1-D contiguous float64 CUDA tensors on ``dev`` (and exactly ``n`` long when n is given): the C ABI takes raw
    pointers and one length, so anything else is an out-of-bounds device access, not an exception."""
    import torch
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float64:
            raise TypeError(f"{what} must be float64 CUDA tensors")
        if t.dim() != 1 or not t.is_contiguous():
            raise ValueError(f"{what} must be 1-D and contiguous")
        if t.device != dev:
            raise ValueError(f"{what} must all live on {dev}")
        if n is not None and t.numel() != n:
            raise ValueError("All evaluation columns must have the same length")


# ------------------------------------------------------------------------------------------------------
_PROGRAM_CACHE: Dict[tuple, CompiledEvaluator] = {}


def _cached_evaluator(exprs: Sequence[Expr], names: Sequence[str]) -> CompiledEvaluator:
    # The reference compiles inside every eval_f64 call (batch.rs:28-30); handles are cached here.
    # keyed on the structural key of each expression (expr.Expr._key), not on id(): ids are only unique while the node
    # is alive, and the cache must not depend on the interning table keeping every node alive forever
    key = (tuple(e._key for e in exprs), tuple(names))
    ev = _PROGRAM_CACHE.get(key)
    if ev is None:
        ev = CompiledEvaluator(list(exprs), list(names))
        if len(_PROGRAM_CACHE) > 256:
            _PROGRAM_CACHE.clear()
        _PROGRAM_CACHE[key] = ev
    return ev


def eval_f64(expressions: Sequence, var_names: Sequence[Sequence[str]], data: Sequence[Sequence[ArrayLike]]):
    """``eval_f64`` (parallel.rs:177-256 / api.rs:352-376).

    ``data[expr][var]`` are columns.  Returns a 2-D ndarray ``(n_exprs, n_points)`` if ANY column is
    an ndarray, else a list of lists.  Expressions that share their variable list and their column
    objects are fused into one multi-output program and evaluated in a single pass."""
    if len(expressions) != len(var_names) or len(expressions) != len(data):
        raise ValueError("exprs, var_names, and data must have the same length")
    exprs = [_as_expr(e) for e in expressions]
    extracted = [[_extract(c) for c in row] for row in data]
    use_numpy = any(flag for row in extracted for _, flag in row)
    if not exprs:
        return np.zeros((0, 0)) if use_numpy else []

    groups: Dict[tuple, List[int]] = {}
    for i, row in enumerate(data):
        key = (tuple(var_names[i]), tuple(id(c) for c in row))
        groups.setdefault(key, []).append(i)
    results: List[Optional[np.ndarray]] = [None] * len(exprs)
    for (names, _), idxs in groups.items():
        cols = [c for c, _ in extracted[idxs[0]]]
        n_points = 1 if not cols else cols[0].size  # batch.rs:18-22
        if n_points == 0:
            for i in idxs:
                results[i] = np.empty(0)
            continue
        ev = _cached_evaluator([exprs[i] for i in idxs], names)
        outs = ev._run_host(cols, n_points, broadcast=False)
        for i, o in zip(idxs, outs):
            results[i] = o
    if use_numpy:
        n_points = results[0].size
        if any(r.size != n_points for r in results):
            raise ValueError("Failed to reshape output: ragged results")
        return np.stack(results) if results else np.zeros((0, 0))
    return [r.tolist() for r in results]


def evaluate_parallel(expressions: Sequence, variables: Sequence[Sequence[str]],
                      values: Sequence[Sequence[Sequence[float]]]):
    """Numeric fast path of ``evaluate_parallel`` (parallel.rs:436-487): every value is a number.

    ``values[expr][var][point]``; a variable with fewer values than the longest one is broadcast from
    its last value (parallel.rs:440-465).  Returns ``list[list[float]]``.  ``None`` (``Value::Skip``)
    and symbolic values belong to the reference's symbolic slow path, which is out of scope here."""
    if len(expressions) != len(variables) or len(expressions) != len(values):
        raise ValueError("expressions, variables and values must have the same length")
    out = []
    for e, names, vals in zip(expressions, variables, values):
        if len(names) != len(vals):
            raise ValueError("EvalColumnMismatch: number of variables and value columns differ")
        for col in vals:
            if any(v is None or isinstance(v, (str, Expr)) for v in col):
                raise NotImplementedError("symbolic / skipped values are outside the numeric fast path")
        n_points = max((len(c) for c in vals), default=1)
        cols = []
        for c in vals:
            if len(c) == 0:
                raise ValueError("empty value column")
            a = np.asarray(c, dtype=np.float64)
            if a.size < n_points:
                a = np.concatenate([a, np.full(n_points - a.size, a[-1])])
            cols.append(a)
        ev = _cached_evaluator([_as_expr(e)], list(names))
        res = ev._run_host(cols, n_points, broadcast=False)[0]
        out.append(res.tolist())
    return out
