"""The C-ABI library on a machine without a GPU: it loads, exports every symbol include/saf_b200.h declares, and the
host-only entry points (program creation, info, export, error reporting) behave; evaluation entry points fail
loudly (SAF_ERR_CUDA) instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from symbanafis_b200 import _lib, workloads
from symbanafis_b200._lib import DeviceProgramHandle, SafError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "saf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"SAF_API\s+[\w\s\*]+?\b(saf_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 17 and "saf_eval_host" in names and "saf_program_export_packed" in names
    L = _lib.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "symbanafis_b200/_lib.py binds a different set than the header declares"


def test_version_and_last_error_are_strings():
    L = _lib.lib()
    assert b"sm_100a" in L.saf_version()
    assert isinstance(L.saf_last_error(), bytes)


def test_host_only_entry_points_work_without_a_device():
    w = workloads.get("clifford")
    p = w.program
    h = DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)
    info = h.info()
    assert (info.param_count, info.n_results) == (2, 2)
    assert info.fp64_slots == 70.0 and info.bytes_per_point == 32.0
    assert "sin" in h.disassemble()
    for ppt in (1, 2, 4, 8):
        pk = h.export_packed(ppt)
        assert pk["double_buffered"] and pk["code_off"] % 32 == 0 and len(pk["image"]) * 8 > pk["code_off"]
    with pytest.raises(SafError) as ei:
        h.export_packed(3)
    assert ei.value.code == _lib.SAF_ERR_INVALID_ARGUMENT


def test_argument_errors_have_distinct_status_codes():
    L = _lib.lib()
    out = C.c_void_p()
    assert L.saf_program_create(None, 0, None, 0, None, 0, 0, 0, None, 0, C.byref(out)) == _lib.SAF_ERR_INVALID_ARGUMENT
    assert L.saf_program_create(None, 0, None, 0, None, 0, 0, 0, None, 0, None) == _lib.SAF_ERR_INVALID_ARGUMENT
    assert L.saf_program_get_info(None, None) == _lib.SAF_ERR_INVALID_ARGUMENT
    assert b"" != L.saf_last_error()
    L.saf_program_destroy(None)  # a no-op, like dropping nothing


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the behaviour WITHOUT a CUDA device")
def test_evaluation_fails_loudly_without_a_device():
    w = workloads.get("single_expr")
    p = w.program
    h = DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)
    x = np.linspace(-1, 1, 16)
    out = np.zeros(16)
    with pytest.raises(SafError) as ei:
        h.eval_host([x], [out], 16)
    assert ei.value.code == _lib.SAF_ERR_CUDA  # no CPU fallback
    # column-length mismatch is diagnosed before any device work (batch.rs:48-50)
    with pytest.raises(SafError) as ei:
        h.eval_host([x[:5]], [out], 16)
    assert ei.value.code == _lib.SAF_ERR_COLUMN_LENGTH_MISMATCH
    # the Python mirror raises too
    from symbanafis_b200.evaluator import CompiledEvaluator
    with pytest.raises(Exception):
        CompiledEvaluator("x^2", ["x"]).eval_batch([x])


def test_program_cache_is_keyed_by_content():
    # batch.rs:28-30: the reference compiles per call; saf_program_intern returns the same handle for the same payload
    import ctypes as C
    from symbanafis_b200 import workloads
    L = _lib.lib()

    def intern(p, consts=None):
        flat = np.ascontiguousarray(p.flat, dtype=np.uint32)
        k = np.ascontiguousarray(p.consts if consts is None else consts, dtype=np.float64)
        pool = np.ascontiguousarray(p.arg_pool, dtype=np.uint32)
        rr = np.ascontiguousarray(p.result_regs, dtype=np.uint32)
        out = C.c_void_p()
        rc = L.saf_program_intern(flat.ctypes.data, flat.size, k.ctypes.data if k.size else None, k.size,
                                  pool.ctypes.data if pool.size else None, pool.size, int(p.param_count),
                                  int(p.workspace_size), rr.ctypes.data, rr.size, C.byref(out))
        assert rc == 0, L.saf_last_error()
        return out.value

    L.saf_program_cache_clear()
    a, b = workloads.get("aizawa").program, workloads.get("clifford").program
    h1, h2, h3 = intern(a), intern(a), intern(b)
    assert h1 == h2 and h1 != h3
    changed = np.array(a.consts, dtype=np.float64).copy()
    changed[0] = -changed[0]
    assert intern(a, changed) not in (h1, h3)  # same bytecode, another constant: another program
    info = _lib.ProgramInfo()
    assert L.saf_program_get_info(C.c_void_p(h1), C.byref(info)) == 0 and info.param_count == 3
    L.saf_program_cache_clear()
    assert intern(a) is not None
    L.saf_program_cache_clear()


def test_program_cache_is_thread_safe():
    import ctypes as C
    import threading
    from symbanafis_b200 import workloads
    L = _lib.lib()
    L.saf_program_cache_clear()
    p = workloads.get("double_pendulum").program
    flat = np.ascontiguousarray(p.flat, dtype=np.uint32)
    k = np.ascontiguousarray(p.consts, dtype=np.float64)
    pool = np.ascontiguousarray(p.arg_pool, dtype=np.uint32)
    rr = np.ascontiguousarray(p.result_regs, dtype=np.uint32)
    got, errs = [None] * 8, []

    def work(i):
        out = C.c_void_p()
        rc = L.saf_program_intern(flat.ctypes.data, flat.size, k.ctypes.data, k.size, pool.ctypes.data if pool.size else None,
                                  pool.size, int(p.param_count), int(p.workspace_size), rr.ctypes.data, rr.size, C.byref(out))
        if rc != 0:
            errs.append(rc)
        got[i] = out.value

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs and len(set(got)) == 1 and got[0]
    L.saf_program_cache_clear()


def test_header_is_valid_c99_and_links_against_the_library(tmp_path):
    # the drop-in boundary is a C ABI: a C (not C++) translation unit that uses the specialisation entry points compiles
    # with -pedantic, links against the library and runs the host-only calls without a device
    import subprocess
    src = tmp_path / "use.c"
    src.write_text(
        '#include <stdio.h>\n#include "saf_b200.h"\n'
        "int main(void) {\n"
        '  const char *exprs[] = {"x*x + sin(x)*exp(-x)"}; const char *params[] = {"x"};\n'
        "  saf_program *p = 0; saf_specialized_info si;\n"
        "  if (saf_compile_exprs(exprs, 1, params, 1, &p) != SAF_OK) return 1;\n"
        "  if (saf_program_specialized_info(p, &si) != SAF_OK || si.compiled != 0) return 2;\n"
        "  if (saf_specialize_available() && (saf_program_specialize(p) != SAF_OK || saf_program_specialized_info(p, &si) != SAF_OK ||\n"
        "      si.compiled != 1 || saf_program_specialized_source(p, 0, 0, 0) == 0)) return 3;\n"
        '  printf("%s\\n", saf_version());\n'
        "  saf_program_destroy(p);\n  return 0;\n}\n")
    exe = tmp_path / "use"
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), _lib.LIB_PATH, f"-Wl,-rpath,{lib_dir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "symbanafis-b200" in r.stdout, (r.returncode, r.stdout, r.stderr)
