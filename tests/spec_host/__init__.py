"""TEST INFRASTRUCTURE ONLY: executes the text of a specialised kernel (symbanafis_b200/csrc/spec.cpp) on the CPU.

The generated CUDA C++ is compiled by g++ against tests/spec_host/shim.h (CUDA vocabulary -> host) together with
driver.cpp (the grid / block / thread loops), so that the CPU test-suite can check what the generator WROTE -- statement
semantics, slot / accumulator handling, tile loop, column rules, in-kernel iteration, prefetch entry, speculative replay
-- against the oracle without a GPU.  Nothing in the product imports this; the product has no CPU path.
"""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "symbanafis_b200", "csrc")
_cache = {}


class HostKernel:
    def __init__(self, source: str):
        h = hashlib.sha1(source.encode())
        for dep in (os.path.join(CSRC, "spec_prelude.cuh"), os.path.join(CSRC, "devmath.cuh"), os.path.join(CSRC, "leanmath.cuh"),
                    os.path.join(CSRC, "saf_isa.h"), os.path.join(HERE, "shim.h"), os.path.join(HERE, "driver.cpp")):
            with open(dep, "rb") as f:
                h.update(f.read())  # a change of any header rebuilds the cached library
        key = h.hexdigest()[:16]
        d = os.path.join(tempfile.gettempdir(), "saf_spec_host")
        os.makedirs(d, exist_ok=True)
        cu, so = os.path.join(d, key + ".cu"), os.path.join(d, key + ".so")
        if not os.path.exists(so):
            with open(cu, "w") as f:
                f.write(source)
            cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-include", os.path.join(HERE, "shim.h"),
                   "-I", CSRC, f'-DSAF_GENERATED="{cu}"', os.path.join(HERE, "driver.cpp"), "-o", so]
            if "saf_spec_pf" in source:
                cmd.insert(1, "-DSAF_HAS_PF")
            subprocess.run(cmd, check=True)
        self.lib = C.CDLL(so)
        self.lib.saf_spec_host_run.argtypes = [C.c_int, C.c_uint, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_ulonglong), C.c_int, C.POINTER(C.c_void_p), C.c_int]
        p, b = C.c_int(), C.c_int()
        self.lib.saf_spec_host_shape(C.byref(p), C.byref(b))
        self.points_per_thread, self.block = p.value, b.value
        self.has_prefetch_entry = "saf_spec_pf" in source

    def run(self, cols, n_points, n_outs, iters=1, grid=3, entry=0, outs=None, col_lens=None):
        cols = [np.ascontiguousarray(c, dtype=np.float64) if c is not None else None for c in cols]
        outs = outs if outs is not None else [np.full(n_points, -777.0) for _ in range(n_outs)]
        cp = (C.c_void_p * 32)(*[c.ctypes.data if c is not None else None for c in cols])
        lens = col_lens if col_lens is not None else [c.size if c is not None else 0 for c in cols]
        cl = (C.c_ulonglong * 32)(*lens)
        op = (C.c_void_p * 32)(*[o.ctypes.data for o in outs])
        rc = self.lib.saf_spec_host_run(entry, grid, n_points, iters, cp, cl, len(cols), op, len(outs))
        assert rc == 0, "entry point missing"
        return outs


def host_kernel(source: str) -> HostKernel:
    k = _cache.get(source)
    if k is None:
        k = _cache[source] = HostKernel(source)
    return k
