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
from .expr import KNOWN_CONSTANTS, Expr, parse, to_string
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
        raw = list(expr) if isinstance(expr, (list, tuple)) else [expr]
        if params is not None and raw and all(isinstance(e, str) for e in raw):
            # expression STRINGS with an explicit parameter list compile in C++ behind the C ABI (saf_compile_exprs,
            # csrc/compile.cpp): no Python on the way from the source text to the device program
            names = [str(p) for p in params]
            try:
                self._h = _lib.DeviceProgramHandle.compile(raw, names)
            except _lib.SafError as e:
                if e.code == _lib.SAF_ERR_INVALID_ARGUMENT:
                    raise ValueError(e.message) from None
                _raise_from_saf(e)
            flat, consts, pool, pc, ws, rr = self._h.reference_payload()
            self.program = Program(flat=flat, consts=consts, arg_pool=pool, param_count=pc, workspace_size=ws,
                                   result_regs=rr, param_names=names)
            self.n_results = len(rr)
            return
        exprs = [_as_expr(e) for e in raw]
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

    def specialize(self):
        """Compiles this evaluator's program into a straight-line sm_100a kernel of its own (saf_program_specialize):
        every later eval_batch / run_chunked / iterate / *_device call launches it instead of the interpreter and
        returns the same bits.  Worth its one-off compilation (tenths of a second; seconds for thousands of
        instructions) when the evaluator is used on many points.  Returns the library's saf_specialized_info."""
        try:
            return self._h.specialize()
        except _lib.SafError as err:
            _raise_from_saf(err)

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


def _format_float(v: float):
    """EvalResult boxing of the reference's Python binding (bindings/python/parallel.rs:118-141): numeric results -- the
    Expr::number / format_float string of parallel.rs:475-484, 680-692 -- reach Python as floats."""
    return float(v)


def _slow_point(expr: Expr, names: Sequence[str], point: Sequence[Optional[float]], was_string: bool):
    """``evaluate_slow_point`` (parallel.rs:630-677) for a point with at least one ``Value::Skip``: numeric variables are
    substituted, skipped ones stay symbolic; the partially evaluated expression comes back as ``Expr`` (input was an
    ``Expr``) or as its string (input was a string).  This is the reference's SYMBOLIC path -- there is nothing to
    evaluate numerically for such a point, on any device.  A maintainer wiring the C ABI into the reference keeps
    calling the reference's own ``evaluate_slow_point`` here."""
    partial = expr.substitute({n: v for n, v in zip(names, point) if v is not None})
    if not partial.variables() and partial.kind == "num":
        return _format_float(partial.value)
    return to_string(partial) if was_string else partial


def evaluate_parallel(expressions: Sequence, variables: Sequence[Sequence[str]],
                      values: Sequence[Sequence[Sequence[Optional[float]]]], symbolic_fallback=None):
    """``evaluate_parallel`` (bindings/python/parallel.rs:34-157 -> drivers/parallel.rs:355-627).

    ``values[expr][var][point]``; a variable with fewer values than the longest one is broadcast from its last value
    (parallel.rs:440-465); ``None`` is ``Value::Skip``.  Per expression:

    * every value numeric -> the numeric fast path (parallel.rs:436-487): one evaluation on the GPU;
    * mixed (parallel.rs:489-613): the points whose values are all numeric are gathered and evaluated on the GPU in ONE
      launch, the others are handed to the symbolic slow path (``symbolic_fallback(expr, names, point_values,
      was_string)``, default: :func:`_slow_point`) -- the compiled evaluator is never asked to run on the CPU.

    Returns ``list[list]``: floats for numeric results, ``Expr`` / ``str`` for symbolic ones (EvalResult boxing of the
    binding, parallel.rs:118-141)."""
    if len(expressions) != len(variables) or len(expressions) != len(values):
        raise ValueError("expressions, variables and values must have the same length")
    slow = symbolic_fallback or _slow_point
    out = []
    for e, names, vals in zip(expressions, variables, values):
        if not isinstance(e, (str, Expr)):
            raise TypeError("expressions must be str or Expr")
        was_string = isinstance(e, str)
        expr = _as_expr(e)
        names = [str(n) for n in names]
        if len(names) != len(vals):
            raise ValueError(f"EvalColumnMismatch: expected {len(names)} value columns, got {len(vals)}")
        cols: List[np.ndarray] = []   # float64, NaN where skipped
        skips: List[Optional[np.ndarray]] = []
        for col in vals:
            if isinstance(col, np.ndarray):
                if col.ndim != 1 or col.dtype != np.float64:
                    raise TypeError("value columns must be 1-D float64 arrays or lists of float / None")
                cols.append(np.ascontiguousarray(col))
                skips.append(None)
                continue
            if any(isinstance(v, (str, Expr)) for v in col):
                raise TypeError("value columns hold floats or None (Value::Expr cannot be built from Python)")
            mask = np.fromiter((v is None for v in col), dtype=bool, count=len(col))
            cols.append(np.asarray([np.nan if v is None else float(v) for v in col], dtype=np.float64))
            skips.append(mask if mask.any() else None)
        n_points = max((c.size for c in cols), default=0)
        if not names:  # zero-variable expression: one value (parallel.rs:399-409)
            ev = _cached_evaluator([expr], [])
            out.append([_format_float(ev._run_host([], 1, broadcast=False)[0][0])])
            continue
        if n_points == 0:
            out.append([])
            continue
        if any(c.size == 0 for c in cols):
            raise ValueError("All evaluation columns must have the same length")  # EvalColumnLengthMismatch (parallel.rs:421-423)
        full, skip_at = [], np.zeros(n_points, dtype=bool)
        for c, m in zip(cols, skips):  # broadcast from the last value, skipped or not
            if c.size < n_points:
                c = np.concatenate([c, np.full(n_points - c.size, c[-1])])
                if m is not None:
                    m = np.concatenate([m, np.full(n_points - m.size, m[-1])])
            full.append(c)
            if m is not None:
                skip_at |= m
        ev = _cached_evaluator([expr], names)
        if not skip_at.any():
            out.append([_format_float(v) for v in ev._run_host(full, n_points, broadcast=False)[0]])
            continue
        row: List = [None] * n_points
        idx = np.flatnonzero(~skip_at)
        if idx.size:  # the numeric points: gathered, ONE launch, scattered back
            res = ev._run_host([np.ascontiguousarray(c[idx]) for c in full], int(idx.size), broadcast=False)[0]
            for i, v in zip(idx, res):
                row[int(i)] = _format_float(v)
        for i in np.flatnonzero(skip_at):
            point = [None if (m is not None and (m[min(i, m.size - 1)])) else float(c[i]) for c, m in zip(full, skips)]
            row[int(i)] = slow(expr, names, point, was_string)
        out.append(row)
    return out
