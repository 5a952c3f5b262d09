"""Specialised kernels (saf_program_specialize; csrc/spec.cpp, spec_prelude.cuh, jit.cpp) against the interpreter and
the oracle.

A specialised kernel is the device program written out as straight-line sm_100a code whose statements are the
interpreter's handlers, compiled at run time by NVRTC.  The bar is therefore the strongest one available: the SAME
BITS as the interpreter on every program and every driver rule (broadcast, missing and short columns, unaligned
buffers, ragged tiles, in-kernel iteration), and -- independently -- the oracle's tolerances of test_gpu_parity.py.
"""
import threading
import zlib

import numpy as np
import pytest

from oracle import OracleProgram
from symbanafis_b200 import workloads
from symbanafis_b200._lib import DeviceProgramHandle, SafError, kernel_launch_count, specialize_available
from symbanafis_b200.compiler import compile_exprs
from symbanafis_b200.expr import parse
from symbanafis_b200.isa import FNOP
from test_gpu_parity import RTOL, ATOL, single_op
from util import FUZZ_BUILTIN1, FUZZ_BUILTIN2, assert_close, bits_equal, random_program

pytestmark = pytest.mark.gpu


def _handle(prog, specialized):
    h = DeviceProgramHandle.create(prog.flat, prog.consts, prog.arg_pool, prog.param_count, prog.workspace_size,
                                   prog.result_regs)
    if specialized:
        info = h.specialize()
        assert info.compiled == 1 and info.cubin_bytes > 0
    return h


def _eval(h, prog, cols, n, broadcast=False):
    outs = [np.full(n, -777.0) for _ in prog.result_regs]
    h.eval_host([np.ascontiguousarray(c, dtype=np.float64) for c in cols], outs, n, broadcast=broadcast)
    return outs


def both(prog, cols, n, broadcast=False):
    """(interpreter outputs, specialised outputs); asserts that the second call really launched the specialised kernel"""
    a = _eval(_handle(prog, False), prog, cols, n, broadcast)
    hs = _handle(prog, True)
    b = _eval(hs, prog, cols, n, broadcast)
    if n:
        si = hs.specialized_info()
        assert si.registers > 0 or si.small_registers > 0, "no specialised kernel was loaded: interpreter fallback"
    return a, b


def assert_same_bits(a, b, what):
    for k, (x, y) in enumerate(zip(a, b)):
        assert bits_equal(x, y), f"{what}: output {k} differs between the interpreter and the specialised kernel"


def test_nvrtc_is_available_on_the_gpu_box():
    assert specialize_available(), "libnvrtc >= 12.8 must be loadable where the GPU tests run"


@pytest.mark.parametrize("name", list(workloads.WORKLOADS))
@pytest.mark.parametrize("variant", ["large", "small"])
@pytest.mark.parametrize("n", [1, 255, 4097, 100_003])
def test_named_programs_same_bits_as_interpreter_and_within_oracle_tolerance(name, n, variant, monkeypatch):
    # every program has two specialised shapes (spec.h): the one for batches that fill the machine and the
    # one-point-per-thread shape for small batches; SAF_SPEC_VARIANT pins the choice the launch would make by size
    monkeypatch.setenv("SAF_SPEC_VARIANT", variant)
    w = workloads.get(name)
    cols = w.make_inputs(n, 23)
    a, b = both(w.program, cols, n)
    assert_same_bits(a, b, f"{name} n={n}")
    ref = OracleProgram.from_program(w.program).run_chunked(cols, n)
    for k, (g, r) in enumerate(zip(b, ref)):
        assert_close(g, r, RTOL, ATOL, what=f"{name} output {k} (specialised vs oracle)")


@pytest.mark.parametrize("name", ["clifford", "aizawa", "double_pendulum"])
@pytest.mark.parametrize("variant", ["large", "small"])
@pytest.mark.parametrize("iters", [1, 2, 7, 50])
def test_in_kernel_iteration_same_bits(name, iters, variant, monkeypatch):
    monkeypatch.setenv("SAF_SPEC_VARIANT", variant)
    w = workloads.get(name)
    n = 3001
    cols = w.make_inputs(n, 5)
    hi, hs = _handle(w.program, False), _handle(w.program, True)
    assert hs.specialized_info().can_iterate == 1
    si = [c.copy() for c in cols]
    ss = [c.copy() for c in cols]
    hi.iterate_host(si, iters)
    before = kernel_launch_count()
    hs.iterate_host(ss, iters)
    assert kernel_launch_count() == before + 1, "iterations must run inside ONE specialised launch"
    assert_same_bits(si, ss, f"{name} x{iters}")
    # and against the oracle stepping the program iters times (short runs: chaotic maps amplify 1-ulp differences)
    if iters <= 2:
        o = OracleProgram.from_program(w.program)
        st = [c.copy() for c in cols]
        for _ in range(iters):
            st = o.eval_batch(st, n)
        for k, (g, r) in enumerate(zip(ss, st)):
            assert_close(g, r, 1e-11, 1e-13, what=f"{name} state {k} after {iters} steps")


@pytest.mark.parametrize("seed", range(24))
def test_random_programs_same_bits(seed):
    rng = np.random.default_rng(5000 + seed)
    prog = random_program(rng, n_params=int(rng.integers(1, 5)), n_instr=int(rng.integers(1, 40)),
                          n_results=int(rng.integers(1, 4)))
    n = 1031
    cols = [rng.uniform(-3, 3, n) for _ in range(prog.param_count)]
    a, b = both(prog, cols, n)
    assert_same_bits(a, b, f"random program {seed}")


SPECIAL = [0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 0.5, np.inf, -np.inf, np.nan, -3.0, 1e-15, 1e5, 5e-324, 1e308]


@pytest.mark.parametrize("op", ["Neg", "Square", "Cube", "Pow4", "Pow3_2", "InvPow3_2", "InvSqrt", "InvSquare", "InvCube",
                                "Recip", "Sqrt", "Copy", "Sin", "Cos", "Exp", "Ln", "RecipExpm1", "ExpSqr", "ExpSqrNeg"])
def test_every_unary_opcode_same_bits(op):
    rng = np.random.default_rng(zlib.crc32(op.encode()))
    x = np.concatenate([rng.uniform(-30, 30, 5000), rng.uniform(-2e6, 2e6, 2000), SPECIAL])
    prog = single_op(op, 1, 0, n_in=1)
    a, b = both(prog, [x], x.size)
    assert_same_bits(a, b, op)


def test_sincos_and_binary_and_ternary_opcodes_same_bits():
    rng = np.random.default_rng(77)
    sp = np.array(SPECIAL)
    x = np.concatenate([rng.normal(0, 100, 5000), np.repeat(sp, sp.size)])
    y = np.concatenate([rng.normal(0, 100, 5000), np.tile(sp, sp.size)])
    z = rng.normal(0, 3, x.size)
    a, b = both(single_op("SinCos", 1, 2, 0, n_in=1, n_out=2), [x], x.size)
    assert_same_bits(a, b, "SinCos")
    for op in ["Add", "Sub", "Mul", "Div", "NegMul", "Pow"]:
        a, b = both(single_op(op, 2, 0, 1, n_in=2), [x, y], x.size)
        assert_same_bits(a, b, op)
    for op in ["MulAdd", "MulSub", "NegMulAdd", "NegMulSub", "Add3", "Mul3"]:
        a, b = both(single_op(op, 3, 0, 1, 2, n_in=3), [x, y, z], x.size)
        assert_same_bits(a, b, op)
    for e in [-7, -1, 0, 1, 2, 3, 10, 31]:
        a, b = both(single_op("Powi", 1, 0, e, n_in=1), [x], x.size)
        assert_same_bits(a, b, f"Powi {e}")


@pytest.mark.parametrize("fn", FUZZ_BUILTIN1)
def test_builtin1_same_bits(fn):
    rng = np.random.default_rng(zlib.crc32(fn.encode()) + 1)
    x = np.concatenate([rng.uniform(-6, 6, 3000), rng.uniform(0.01, 40, 1500), SPECIAL])
    a, b = both(single_op("Builtin1", 1, FNOP[fn], 0, n_in=1), [x], x.size)
    assert_same_bits(a, b, fn)


@pytest.mark.parametrize("fn", FUZZ_BUILTIN2 + ["zeta_deriv"])
def test_builtin2_same_bits(fn):
    rng = np.random.default_rng(zlib.crc32(fn.encode()) + 2)
    n = 1500
    if fn in ("bessel_j", "bessel_y", "bessel_i", "bessel_k", "hermite", "polygamma", "zeta_deriv"):
        u = rng.integers(-3 if fn.startswith("bessel") else 0, 6, n).astype(float)
        v = rng.uniform(0.05, 25, n)
    else:
        u, v = rng.normal(0, 5, n), rng.normal(0, 5, n)
    a, b = both(single_op("Builtin2", 2, FNOP[fn], 0, 1, n_in=2), [u, v], n)
    assert_same_bits(a, b, fn)


def test_builtin3_and_4_same_bits():
    rng = np.random.default_rng(8)
    n = 1200
    l = rng.integers(0, 7, n).astype(float)
    m = np.array([rng.integers(-int(li), int(li) + 1) for li in l], dtype=float)
    x, theta, phi = rng.uniform(-1, 1, n), rng.uniform(0, 3.1, n), rng.uniform(-3, 3, n)
    a, b = both(single_op("Builtin3", 3, FNOP["assoc_legendre"], 0, 1, 2, n_in=3), [l, m, x], n)
    assert_same_bits(a, b, "assoc_legendre")
    a, b = both(single_op("Builtin4", 4, FNOP["spherical_harmonic"], 0, 1, 2, 3, n_in=4), [l, m, theta, phi], n)
    assert_same_bits(a, b, "spherical_harmonic")


def test_compiled_expressions_with_prefix_fusion_same_bits():
    # affine prefixes (sin(a*x), sincos(f*x + p), exp(-u/w)), fused chains and constant results
    exprs = ["sin(2.5*x) + cos(2.5*x)", "exp(-(x - 0.3)^2 / 0.7) * sin(1.5*x + 0.25)", "1/(exp(0.5*y + 1) - 1)", "3.25",
             "x", "x*y/(1 + x^2) - sqrt(abs(y))", "ln(x^2 + 1) + x^(3/2)"]
    rng = np.random.default_rng(31)
    n = 4099
    x, y = rng.uniform(0.01, 4, n), rng.uniform(-4, 4, n)
    prog = compile_exprs([parse(e) for e in exprs], ["x", "y"])
    a, b = both(prog, [x, y], n)
    assert_same_bits(a, b, "compiled expressions")
    ref = OracleProgram.from_program(prog).eval_batch([x, y], n)
    for k, (g, r) in enumerate(zip(b, ref)):
        assert_close(g, r, RTOL, ATOL, what=f"expression {k}")


def test_driver_rules_broadcast_missing_short_unaligned_empty():
    prog = single_op("MulAdd", 3, 0, 1, 2, n_in=3)
    rng = np.random.default_rng(9)
    n = 5000
    p, q, r = rng.normal(size=n), rng.normal(size=1700), rng.normal(size=1)
    o = OracleProgram.from_program(prog)
    hs = _handle(prog, True)
    assert bits_equal(_eval(hs, prog, [p, q, r], n, broadcast=True)[0], o.eval_batch([p, q, r], n)[0])
    assert bits_equal(_eval(hs, prog, [p, p], n)[0], o.eval_batch([p, p], n)[0])  # missing column reads 0.0
    assert bits_equal(_eval(hs, prog, [p] * 5, n)[0], o.eval_batch([p] * 5, n)[0])  # extra columns ignored
    assert bits_equal(_eval(hs, prog, [], 3)[0], np.zeros(3))
    hs.eval_host([np.empty(0)] * 3, [np.empty(0)], 0)  # n_points == 0 is Ok
    with pytest.raises(SafError) as ei:
        hs.eval_host([p, q, r], [np.empty(n)], n)
    assert ei.value.code == 1
    # 8-byte aligned, not 16: the scalar load / store path of a four-points-per-thread kernel
    w = workloads.get("single_expr")
    hw = _handle(w.program, True)
    assert hw.specialized_info().points_per_thread == 4
    m = 700_001  # large-batch shape (>= 12 warps per SM at four points per thread)
    x = w.make_inputs(m + 1, 1)[0]
    out = np.empty(m + 1)
    hw.eval_host([x[1:]], [out[1:]], m)
    assert hw.specialized_info().registers > 0
    ref = np.empty(m)
    _handle(w.program, False).eval_host([x[1:].copy()], [ref], m)
    assert bits_equal(out[1:], ref)


def test_large_batch_several_tiles_per_block_and_host_pipeline():
    w = workloads.get("aizawa")
    n = (1 << 21) + 4567  # two pipeline chunks of the host path; far more tiles than resident blocks
    cols = w.make_inputs(n, 3)
    a, b = both(w.program, cols, n)
    assert_same_bits(a, b, "aizawa large")


def test_shape_follows_the_batch_size():
    w = workloads.get("double_pendulum")
    h = _handle(w.program, True)
    si = h.specialized_info()
    assert si.compiled == 1 and si.small_compiled == 0 and si.points_per_thread == 2
    n = 50_000  # 50 k pendulums on 148 SMs: 5 warps per SM at two points per thread -> the small-batch shape
    st = w.make_inputs(n, 1)
    h.iterate_host(st, 3)
    si = h.specialized_info()
    assert si.small_compiled == 1 and si.small_points_per_thread == 1 and si.small_registers > 0 and si.registers == -1
    n = 1_000_000
    st = w.make_inputs(n, 1)
    h.iterate_host(st, 1)
    assert h.specialized_info().registers > 0
    assert "SPEC_P 1" in h.specialized_source(small_batch=True) and "SPEC_P 2" in h.specialized_source()


def test_long_program_runs_in_step_behind_block_barriers():
    w = workloads.get("terrain_gradient")
    h = _handle(w.program, True)
    si = h.specialized_info()
    src = h.specialized_source()
    assert si.block_threads == 512 and src.count("__syncthreads();") >= 5 and "#define SPEC_OUTLINE 1" in src
    n = 300_000  # > 12 warps per SM at four points per thread: the large shape, a ragged last tile
    cols = w.make_inputs(n, 6)
    a = _eval(_handle(w.program, False), w.program, cols, n)
    b = _eval(h, w.program, cols, n)
    assert h.specialized_info().registers > 0
    assert_same_bits(a, b, "terrain large shape")


def test_device_resident_in_place_iteration_same_bits():
    torch = pytest.importorskip("torch")
    w = workloads.get("clifford")
    n = 200_003
    cols = w.make_inputs(n, 11)
    res = []
    for spec in (False, True):
        h = _handle(w.program, spec)
        st = [torch.from_numpy(c.copy()).cuda() for c in cols]
        h.iterate_device([t.data_ptr() for t in st], n, 25, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        res.append([t.cpu().numpy() for t in st])
    assert_same_bits(res[0], res[1], "device-resident clifford x25")


def test_single_passes_launch_the_prefetch_entry_also_in_place():
    torch = pytest.importorskip("torch")
    w = workloads.get("clifford")
    n = 1_000_003  # several tiles per block: every block prefetches its next tile while it evaluates the current one
    cols = w.make_inputs(n, 4)
    h = _handle(w.program, True)
    s = torch.cuda.current_stream().cuda_stream
    xs = [torch.from_numpy(c.copy()).cuda() for c in cols]
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in cols]
    h.eval_device([t.data_ptr() for t in xs], [n, n], [o.data_ptr() for o in outs], n, stream=s)
    si = h.specialized_info()
    assert si.has_prefetch_entry == 1 and si.prefetch_registers > 0
    # in place: outputs alias the columns (what one step of saf_iterate_device does); tiles never overlap
    h.eval_device([t.data_ptr() for t in xs], [n, n], [t.data_ptr() for t in xs], n, stream=s)
    torch.cuda.synchronize()
    ref = _eval(_handle(w.program, False), w.program, cols, n)
    assert_same_bits(ref, [o.cpu().numpy() for o in outs], "clifford single pass, prefetch entry")
    assert_same_bits(ref, [t.cpu().numpy() for t in xs], "clifford single pass in place, prefetch entry")


def test_program_too_long_to_specialise_keeps_running_on_the_interpreter(monkeypatch):
    monkeypatch.setenv("SAF_SPEC_MAX_INSTRUCTIONS", "10")  # the shipped limit is 20000 device instructions
    w = workloads.get("aizawa")
    h = _handle(w.program, False)
    with pytest.raises(SafError) as ei:
        h.specialize()
    assert ei.value.code == 6 and "interpreted" in ei.value.message
    n = 257
    cols = w.make_inputs(n, 4)
    got = _eval(h, w.program, cols, n)
    ref = OracleProgram.from_program(w.program).eval_batch(cols, n)
    for g, r in zip(got, ref):
        assert_close(g, r, RTOL, ATOL, what="refused program on the interpreter")


def test_one_specialised_handle_from_many_host_threads():
    w = workloads.get("aizawa")
    h = _handle(w.program, True)
    n = 40_000
    cols = w.make_inputs(n, 2)
    ref = _eval(_handle(w.program, False), w.program, cols, n)
    errs = []

    def work():
        try:
            for _ in range(5):
                got = _eval(h, w.program, cols, n)
                for g, r in zip(got, ref):
                    assert bits_equal(g, r)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work) for _ in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[0]


def test_speculative_body_replays_carefully_outside_the_fast_range():
    # arguments beyond 2^20 (trig) and 700 (exp), Inf and NaN, mixed with ordinary ones in the same thread / warp: the
    # speculative copy of the body records them and the careful copy recomputes the iteration -- same bits as the
    # interpreter, whose leaves branch to the math library per call
    rng = np.random.default_rng(12)
    n = 20_011
    x = rng.uniform(-3, 3, n)
    y = rng.uniform(-3, 3, n)
    idx = rng.choice(n, 400, replace=False)
    x[idx[:100]] = rng.uniform(1.1e6, 1e12, 100)
    y[idx[100:200]] = -rng.uniform(1.1e6, 1e15, 100)
    x[idx[200:250]] = np.inf
    y[idx[250:300]] = np.nan
    x[idx[300:400]] = rng.uniform(700, 800, 100)
    prog = compile_exprs([parse("sin(1.3*x) + 0.7*cos(0.9*y)"), parse("exp(-x) * sin(y) + cos(x)")], ["x", "y"])
    for variant in ("large", "small"):
        import os
        os.environ["SAF_SPEC_VARIANT"] = variant
        try:
            a, b = both(prog, [x, y], n)
        finally:
            del os.environ["SAF_SPEC_VARIANT"]
        assert_same_bits(a, b, f"slow paths ({variant})")
    w = workloads.get("clifford")  # iterated: the replay restores the state of the iteration it repeats
    st0 = w.make_inputs(n, 3)
    st0[0][idx[:50]] = 3e7
    hi, hs = _handle(w.program, False), _handle(w.program, True)
    si_, ss_ = [c.copy() for c in st0], [c.copy() for c in st0]
    hi.iterate_host(si_, 5)
    hs.iterate_host(ss_, 5)
    assert_same_bits(si_, ss_, "clifford with out-of-range start")


def test_specialised_source_is_inspectable():
    w = workloads.get("clifford")
    h = _handle(w.program, True)
    src = h.specialized_source()
    assert "saf_spec" in src and "lm::" not in src.split("spec_prelude.cuh")[0]
    assert "sp::sin_v" in src and "sp::sin_f" in src and "__fma_rn" in src
    si = h.specialized_info()
    assert si.compile_ms > 0 and si.source_bytes == len(src)


def test_gpu_execution_equals_cpu_execution_of_the_same_generated_text():
    # tests/spec_host compiles the generated text with g++ (CUDA vocabulary shimmed onto IEEE operations and libm's fma):
    # inside the fast range of the lean sin / cos / exp kernels every statement is an exactly rounded operation, so the
    # B200 and the CPU must produce the same bits from the same text
    from spec_host import host_kernel
    for name, iters in (("clifford", 3), ("aizawa", 5), ("single_expr", 1)):
        w = workloads.get(name)
        n = 4099
        cols = w.make_inputs(n, 8)
        h = _handle(w.program, True)
        k = host_kernel(h.specialized_source())
        n_out = len(w.program.result_regs)
        if iters == 1:
            gpu = _eval(h, w.program, cols, n)
            cpu = k.run(cols, n, n_out, grid=2)
        else:
            gpu = [c.copy() for c in cols]
            h.iterate_host(gpu, iters)
            cpu = [c.copy() for c in cols]
            k.run(cpu, n, n_out, iters=iters, grid=2, outs=cpu)
        assert_same_bits(gpu, cpu, f"{name}: GPU vs CPU execution of the generated kernel")


def test_auto_mode_specialises_a_program_once_it_has_worked_enough():
    # SAF_SPECIALIZE is read once per process: a child process runs with auto
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np\n"
        "from symbanafis_b200 import workloads\n"
        "from symbanafis_b200._lib import DeviceProgramHandle\n"
        "w = workloads.get('aizawa'); p = w.program\n"
        "h = DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)\n"
        "n = 400_000; st = w.make_inputs(n, 1)\n"
        "h.iterate_host([c.copy() for c in st], 100)\n"        # 25 instr x 4e5 x 100 = 1e9 < 2e10: interpreter
        "assert h.specialized_info().compiled == 0\n"
        "a = [c.copy() for c in st]; h.iterate_host(a, 3000)\n"  # 3e10: still the interpreter for THIS call, counted
        "b = [c.copy() for c in st]; h.iterate_host(b, 3000)\n"  # specialised at this launch
        "si = h.specialized_info(); assert si.compiled == 1 and max(si.registers, si.small_registers) > 0, (si.compiled, si.registers)\n"
        "assert all(np.array_equal(x.view(np.uint64), y.view(np.uint64)) for x, y in zip(a, b))\n"
        "print('auto ok')\n")
    env = dict(os.environ, SAF_SPECIALIZE="auto")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "auto ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("name", list(workloads.WORKLOADS))
@pytest.mark.parametrize("n", [1, 777, 100_003])
def test_no_write_outside_the_outputs_and_no_write_to_the_columns(name, n):
    # compute-sanitizer is closed on the GPU pool; the specialised kernels address global memory only through
    # load_col / store_out (spec_prelude.cuh), whose bounds this checks from outside: 64-element guard bands around every
    # output stay untouched, the columns come back unchanged, ragged and 8-byte-aligned placements included
    torch = pytest.importorskip("torch")
    w = workloads.get(name)
    cols = w.make_inputs(n, 13)
    n_out = len(w.program.result_regs)
    G = 64
    s = torch.cuda.current_stream().cuda_stream
    for machine in ("specialized", "interpreter"):
        h = _handle(w.program, machine == "specialized")
        for shift in (0, 1):  # shift 1: 8-byte aligned, not 16 -> the scalar access paths
            d_cols = [torch.from_numpy(c.copy()).cuda() for c in cols]
            arena = [torch.full((n + 2 * G + 1,), -12345.678, dtype=torch.float64, device="cuda") for _ in range(n_out)]
            outs = [a[G + shift:G + shift + n] for a in arena]
            h.eval_device([t.data_ptr() for t in d_cols], [n] * len(d_cols), [o.data_ptr() for o in outs], n, stream=s)
            torch.cuda.synchronize()
            for a in arena:
                assert bool((a[:G + shift] == -12345.678).all()) and bool((a[G + shift + n:] == -12345.678).all()), \
                    f"{name} {machine}: write outside the output column (n={n}, shift={shift})"
            for t, c in zip(d_cols, cols):
                assert np.array_equal(t.cpu().numpy().view(np.uint64), c.view(np.uint64)), f"{name} {machine}: a column was modified"


def test_rolled_long_sum_same_bits_on_the_gpu(monkeypatch):
    # csrc/reroll.inc (opt-in): the 110-term gradient as a loop over its terms -- slower than the straight-line kernel
    #  but the same arithmetic: the same bits
    w = workloads.get("terrain_gradient")
    n = 150_001
    cols = w.make_inputs(n, 21)
    ref = _eval(_handle(w.program, False), w.program, cols, n)
    monkeypatch.setenv("SAF_SPEC_REROLL", "1")
    h = _handle(w.program, True)
    assert "saf_rr_kt" in h.specialized_source()
    got = _eval(h, w.program, cols, n)
    assert max(h.specialized_info().registers, h.specialized_info().small_registers) > 0
    assert_same_bits(ref, got, "rolled terrain gradient")
