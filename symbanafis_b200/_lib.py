"""ctypes binding of ``libsaf_b200.so`` (C ABI in ``include/saf_b200.h``).

There is no Python or CPU fallback: if the shared library is missing, or no CUDA device is
present, evaluation raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SAF_B200_LIB") or os.path.join(_PKG, "libsaf_b200.so")  # env: tuning builds only
CSRC = os.path.join(_PKG, "csrc")

SAF_OK = 0
SAF_ERR_COLUMN_LENGTH_MISMATCH = 1
SAF_ERR_INVALID_PROGRAM = 2
SAF_ERR_ALLOC = 3
SAF_ERR_CUDA = 4
SAF_ERR_INVALID_ARGUMENT = 5
SAF_ERR_UNSUPPORTED = 6
SAF_EVAL_BROADCAST = 1


class SafError(RuntimeError):
    """A non-zero status from the C ABI; ``code`` is one of the ``SAF_ERR_*`` values."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[saf {code}] {message}")
        self.code = code
        self.message = message


class ProgramInfo(C.Structure):
    _fields_ = [
        ("param_count", C.c_uint32), ("n_results", C.c_uint32), ("n_ref_instructions", C.c_uint32),
        ("n_dev_instructions", C.c_uint32), ("n_slots", C.c_uint32), ("n_consts", C.c_uint32),
        ("n_acc_forwards", C.c_uint32), ("quirk_flags", C.c_uint32),
        ("fp64_slots", C.c_double), ("bytes_per_point", C.c_double),
    ]


class SpecializedInfo(C.Structure):
    _fields_ = [
        ("compiled", C.c_int), ("can_iterate", C.c_int), ("points_per_thread", C.c_int), ("block_threads", C.c_int),
        ("min_blocks", C.c_int), ("registers", C.c_int), ("local_bytes", C.c_int), ("resident_blocks", C.c_int),
        ("generate_ms", C.c_double), ("compile_ms", C.c_double), ("cubin_bytes", C.c_size_t), ("source_bytes", C.c_size_t),
        ("small_compiled", C.c_int), ("small_points_per_thread", C.c_int), ("small_block_threads", C.c_int),
        ("small_registers", C.c_int), ("small_compile_ms", C.c_double),
        ("has_prefetch_entry", C.c_int), ("prefetch_registers", C.c_int),
    ]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/`` with nvcc (``-gencode arch=compute_100a,code=sm_100a -lineinfo``)."""
    args = ["make", "-C", CSRC]
    if force:
        args.append("-B")
    proc = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("building libsaf_b200.so failed")
    return LIB_PATH


_lib: Optional[C.CDLL] = None

EXPORTED_SYMBOLS = [
    "saf_program_create", "saf_program_fuse", "saf_program_destroy", "saf_program_intern", "saf_program_cache_clear",
    "saf_program_get_info",
    "saf_program_disassemble", "saf_program_export", "saf_program_export_packed", "saf_eval_device", "saf_eval_host",
    "saf_eval_host_multi", "saf_iterate_device", "saf_iterate_host", "saf_last_error",
    "saf_device_count", "saf_kernel_launch_count", "saf_version", "saf_measure_fp64_peak",
    "saf_measure_clifford_probe", "saf_compile_exprs", "saf_program_reference_payload",
    "saf_program_specialize", "saf_program_specialized_info", "saf_program_specialized_source", "saf_specialize_available",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, u32, u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64
    PP = C.POINTER(C.c_void_p)
    L.saf_program_create.argtypes = [vp, sz, vp, sz, vp, sz, u32, u32, vp, u32, PP]
    L.saf_program_create.restype = C.c_int
    L.saf_program_intern.argtypes = [vp, sz, vp, sz, vp, sz, u32, u32, vp, u32, PP]
    L.saf_program_intern.restype = C.c_int
    L.saf_program_cache_clear.argtypes = []
    L.saf_program_cache_clear.restype = None
    L.saf_program_fuse.argtypes = [PP, sz, PP]
    L.saf_program_fuse.restype = C.c_int
    L.saf_compile_exprs.argtypes = [C.POINTER(C.c_char_p), sz, C.POINTER(C.c_char_p), sz, PP]
    L.saf_compile_exprs.restype = C.c_int
    L.saf_program_reference_payload.argtypes = [vp, sz, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz),
                                                C.POINTER(u32), C.POINTER(u32), vp, sz, C.POINTER(sz)]
    L.saf_program_reference_payload.restype = C.c_int
    L.saf_program_destroy.argtypes = [vp]
    L.saf_program_destroy.restype = None
    L.saf_program_get_info.argtypes = [vp, C.POINTER(ProgramInfo)]
    L.saf_program_get_info.restype = C.c_int
    L.saf_program_disassemble.argtypes = [vp, C.c_char_p, sz]
    L.saf_program_disassemble.restype = sz
    L.saf_program_export.argtypes = [vp, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz), vp, vp]
    L.saf_program_export.restype = C.c_int
    L.saf_program_export_packed.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz), C.POINTER(u32), C.POINTER(u32),
                                            vp, vp, vp, vp, vp, C.POINTER(C.c_int)]
    L.saf_program_export_packed.restype = C.c_int
    L.saf_eval_device.argtypes = [vp, PP, C.POINTER(sz), sz, PP, sz, C.c_uint, vp]
    L.saf_eval_device.restype = C.c_int
    L.saf_eval_host.argtypes = [vp, PP, C.POINTER(sz), sz, PP, sz, C.c_uint]
    L.saf_eval_host.restype = C.c_int
    L.saf_eval_host_multi.argtypes = [vp, PP, C.POINTER(sz), sz, PP, sz, C.c_uint, C.POINTER(C.c_int), C.c_int]
    L.saf_eval_host_multi.restype = C.c_int
    L.saf_iterate_device.argtypes = [vp, PP, sz, sz, u64, vp]
    L.saf_iterate_device.restype = C.c_int
    L.saf_iterate_host.argtypes = [vp, PP, sz, sz, u64]
    L.saf_iterate_host.restype = C.c_int
    L.saf_last_error.restype = C.c_char_p
    L.saf_device_count.argtypes = [C.POINTER(C.c_int)]
    L.saf_device_count.restype = C.c_int
    L.saf_kernel_launch_count.restype = u64
    L.saf_measure_fp64_peak.argtypes = [C.POINTER(C.c_double)]
    L.saf_measure_fp64_peak.restype = C.c_int
    L.saf_measure_clifford_probe.argtypes = [sz, C.c_int, C.POINTER(C.c_double)]
    L.saf_measure_clifford_probe.restype = C.c_int
    L.saf_version.restype = C.c_char_p
    L.saf_program_specialize.argtypes = [vp]
    L.saf_program_specialize.restype = C.c_int
    L.saf_program_specialized_info.argtypes = [vp, C.POINTER(SpecializedInfo)]
    L.saf_program_specialized_info.restype = C.c_int
    L.saf_program_specialized_source.argtypes = [vp, C.c_int, C.c_char_p, sz]
    L.saf_program_specialized_source.restype = sz
    L.saf_specialize_available.argtypes = []
    L.saf_specialize_available.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != SAF_OK:
        raise SafError(rc, lib().saf_last_error().decode(errors="replace"))


def ptr_array(ptrs: Sequence[int]):
    a = (C.c_void_p * max(1, len(ptrs)))()
    for i, p in enumerate(ptrs):
        a[i] = p
    return a


def size_array(vals: Sequence[int]):
    return (C.c_size_t * max(1, len(vals)))(*vals)


def specialize_available() -> bool:
    """Can specialised kernels be compiled in this process (NVRTC >= 12.8 loadable)?"""
    return bool(lib().saf_specialize_available())


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().saf_device_count(C.byref(n))
    return n.value if rc == SAF_OK else 0


def measure_fp64_peak() -> float:
    """FP64 pipe peak of the current device in DFMA lane-instructions per second."""
    v = C.c_double(0.0)
    check(lib().saf_measure_fp64_peak(C.byref(v)))
    return v.value


def measure_clifford_probe(n_points: int, iters: int) -> float:
    """Seconds of the hard-coded Clifford kernel (no interpreter) for n_points x iters: the ideal-kernel reference."""
    v = C.c_double(0.0)
    check(lib().saf_measure_clifford_probe(n_points, iters, C.byref(v)))
    return v.value


def kernel_launch_count() -> int:
    return int(lib().saf_kernel_launch_count())


class DeviceProgramHandle:
    """Owns a ``saf_program*``."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    @classmethod
    def create(cls, flat, consts, arg_pool, param_count: int, workspace_size: int,
               result_regs: Sequence[int]) -> "DeviceProgramHandle":
        flat = np.ascontiguousarray(flat, dtype=np.uint32)
        consts = np.ascontiguousarray(consts, dtype=np.float64)
        arg_pool = np.ascontiguousarray(arg_pool, dtype=np.uint32)
        rr = np.ascontiguousarray(result_regs, dtype=np.uint32)
        out = C.c_void_p()
        check(lib().saf_program_create(
            flat.ctypes.data, flat.size, consts.ctypes.data if consts.size else None, consts.size,
            arg_pool.ctypes.data if arg_pool.size else None, arg_pool.size,
            int(param_count), int(workspace_size), rr.ctypes.data, rr.size, C.byref(out)))
        return cls(out.value)

    @classmethod
    def compile(cls, exprs: Sequence[str], params: Sequence[str]) -> "DeviceProgramHandle":
        """Expression strings -> one multi-output program, compiled by the C++ compiler behind the C ABI
        (``saf_compile_exprs``, csrc/compile.cpp)."""
        ea = (C.c_char_p * max(1, len(exprs)))(*[e.encode() for e in exprs])
        pa = (C.c_char_p * max(1, len(params)))(*[p.encode() for p in params])
        out = C.c_void_p()
        check(lib().saf_compile_exprs(ea, len(exprs), pa, len(params), C.byref(out)))
        return cls(out.value)

    def reference_payload(self, index: int = 0):
        """(flat uint32[], consts f64[], arg_pool uint32[], param_count, workspace_size, result_regs) of component `index`."""
        nw, nk, npool, nr = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        pc, ws = C.c_uint32(), C.c_uint32()
        check(lib().saf_program_reference_payload(self._h, index, None, 0, C.byref(nw), None, 0, C.byref(nk), None, 0,
                                                  C.byref(npool), C.byref(pc), C.byref(ws), None, 0, C.byref(nr)))
        flat = np.zeros(max(1, nw.value), dtype=np.uint32)
        consts = np.zeros(max(1, nk.value), dtype=np.float64)
        pool = np.zeros(max(1, npool.value), dtype=np.uint32)
        rr = np.zeros(max(1, nr.value), dtype=np.uint32)
        check(lib().saf_program_reference_payload(self._h, index, flat.ctypes.data, flat.size, C.byref(nw), consts.ctypes.data,
                                                  consts.size, C.byref(nk), pool.ctypes.data, pool.size, C.byref(npool),
                                                  C.byref(pc), C.byref(ws), rr.ctypes.data, rr.size, C.byref(nr)))
        return (flat[:nw.value], consts[:nk.value], pool[:npool.value], int(pc.value), int(ws.value),
                [int(r) for r in rr[:nr.value]])

    @classmethod
    def fuse(cls, handles: Sequence["DeviceProgramHandle"]) -> "DeviceProgramHandle":
        arr = ptr_array([h._h.value for h in handles])
        out = C.c_void_p()
        check(lib().saf_program_fuse(arr, len(handles), C.byref(out)))
        return cls(out.value)

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.saf_program_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def ptr(self) -> C.c_void_p:
        return self._h

    def info(self) -> ProgramInfo:
        info = ProgramInfo()
        check(lib().saf_program_get_info(self._h, C.byref(info)))
        return info

    def specialize(self) -> "SpecializedInfo":
        """Compiles this program into a straight-line sm_100a kernel of its own (saf_program_specialize): every later
        eval_* / iterate_* of this handle launches it instead of the interpreter.  Same bits, no dispatch."""
        check(lib().saf_program_specialize(self._h))
        return self.specialized_info()

    def specialized_info(self) -> "SpecializedInfo":
        info = SpecializedInfo()
        check(lib().saf_program_specialized_info(self._h, C.byref(info)))
        return info

    def specialized_source(self, small_batch: bool = False) -> str:
        n = lib().saf_program_specialized_source(self._h, int(small_batch), None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().saf_program_specialized_source(self._h, int(small_batch), buf, n + 1)
        return buf.value.decode()

    def disassemble(self) -> str:
        n = lib().saf_program_disassemble(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().saf_program_disassemble(self._h, buf, n + 1)
        return buf.value.decode()

    def export(self):
        """(code uint64[], consts f64[], param_slot uint16[], result_field uint32[])"""
        nc, nk = C.c_size_t(), C.c_size_t()
        check(lib().saf_program_export(self._h, None, 0, C.byref(nc), None, 0, C.byref(nk), None, None))
        info = self.info()
        code = np.zeros(nc.value, dtype=np.uint64)
        consts = np.zeros(max(1, nk.value), dtype=np.float64)
        pslot = np.zeros(max(1, info.param_count), dtype=np.uint16)
        rfield = np.zeros(max(1, info.n_results), dtype=np.uint32)
        check(lib().saf_program_export(self._h, code.ctypes.data, code.size, C.byref(nc), consts.ctypes.data,
                                       consts.size, C.byref(nk), pslot.ctypes.data, rfield.ctypes.data))
        return code, consts[:nk.value], pslot[:info.param_count], rfield[:info.n_results]

    def export_packed(self, points_per_thread: int):
        """The packed micro-instruction image for one register-file layout (csrc/uop.h):
        dict(image uint64[], code_off, n_slots, x_off int32[3] = staging x0, x1 and the ACC slot, param_off int32[], result_off int32[],
        result_is_k uint8[], result_off_even int32[], double_buffered)."""
        nw, co, ns = C.c_size_t(), C.c_uint32(), C.c_uint32()
        check(lib().saf_program_export_packed(self._h, points_per_thread, None, 0, C.byref(nw), C.byref(co),
                                              C.byref(ns), None, None, None, None, None, None))
        info = self.info()
        image = np.zeros(nw.value, dtype=np.uint64)
        x_off = np.zeros(3, dtype=np.int32)  # x0, x1, acc
        poff = np.zeros(max(1, info.param_count), dtype=np.int32)
        roff = np.zeros(max(1, info.n_results), dtype=np.int32)
        rk = np.zeros(max(1, info.n_results), dtype=np.uint8)
        roff_even = np.zeros(max(1, info.n_results), dtype=np.int32)
        dbuf = C.c_int(0)
        check(lib().saf_program_export_packed(self._h, points_per_thread, image.ctypes.data, image.size, C.byref(nw),
                                              C.byref(co), C.byref(ns), x_off.ctypes.data, poff.ctypes.data,
                                              roff.ctypes.data, rk.ctypes.data, roff_even.ctypes.data, C.byref(dbuf)))
        return {"image": image, "code_off": co.value, "n_slots": ns.value, "x_off": x_off,
                "param_off": poff[:info.param_count], "result_off": roff[:info.n_results],
                "result_is_k": rk[:info.n_results],
                "points_per_thread": points_per_thread // 1000 if points_per_thread > 1000 else points_per_thread,
                "block_threads": points_per_thread % 1000 if points_per_thread > 1000 else 128,
                "result_off_even": roff_even[:info.n_results], "double_buffered": bool(dbuf.value)}

    # -- evaluation ----------------------------------------------------------------------------
    def eval_host(self, cols: Sequence[np.ndarray], outs: Sequence[np.ndarray], n_points: int,
                  broadcast: bool = False, devices: Optional[Sequence[int]] = None) -> None:
        for a in list(cols) + list(outs):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.ndim == 1 and a.flags.c_contiguous):
                raise ValueError("eval_host: columns and outputs must be 1-D contiguous float64 arrays")
        if any(o.size < n_points or not o.flags.writeable for o in outs):
            raise ValueError("eval_host: every output array must be writable and hold n_points values")
        cp = ptr_array([c.ctypes.data for c in cols])
        cl = size_array([c.size for c in cols])
        op = ptr_array([o.ctypes.data for o in outs])
        flags = SAF_EVAL_BROADCAST if broadcast else 0
        if devices is None:
            check(lib().saf_eval_host(self._h, cp, cl, len(cols), op, n_points, flags))
        else:
            dv = (C.c_int * len(devices))(*devices)
            check(lib().saf_eval_host_multi(self._h, cp, cl, len(cols), op, n_points, flags, dv, len(devices)))

    def eval_device(self, col_ptrs: Sequence[int], col_lens: Sequence[int], out_ptrs: Sequence[int],
                    n_points: int, broadcast: bool = False, stream: int = 0) -> None:
        flags = SAF_EVAL_BROADCAST if broadcast else 0
        check(lib().saf_eval_device(self._h, ptr_array(col_ptrs), size_array(col_lens), len(col_ptrs),
                                    ptr_array(out_ptrs), n_points, flags, C.c_void_p(stream)))

    def iterate_host(self, state: Sequence[np.ndarray], iters: int) -> None:
        n = state[0].size if state else 0
        for s in state:  # one length goes to C for every pointer
            if not (isinstance(s, np.ndarray) and s.dtype == np.float64 and s.ndim == 1 and s.flags.c_contiguous
                    and s.flags.writeable and s.size == n):
                raise ValueError("iterate_host: state columns must be writable 1-D contiguous float64 arrays of one length")
        check(lib().saf_iterate_host(self._h, ptr_array([s.ctypes.data for s in state]), len(state), n, iters))

    def iterate_device(self, state_ptrs: Sequence[int], n_points: int, iters: int, stream: int = 0) -> None:
        check(lib().saf_iterate_device(self._h, ptr_array(state_ptrs), len(state_ptrs), n_points, iters,
                                       C.c_void_p(stream)))
