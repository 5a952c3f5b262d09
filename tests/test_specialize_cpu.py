"""Host half of the specialised kernels (csrc/spec.cpp, jit.cpp), no GPU: the generated text and the NVRTC compilation.

NVRTC is a compiler: it runs without a device, so the whole path up to the cubin is checked here; what the cubin
computes is checked on the GPU (tests/test_gpu_specialized.py: same bits as the interpreter)."""
import re

import numpy as np
import pytest

from symbanafis_b200 import workloads
from symbanafis_b200._lib import DeviceProgramHandle, SafError, specialize_available
from symbanafis_b200.compiler import compile_exprs
from symbanafis_b200.expr import parse

needs_nvrtc = pytest.mark.skipif(not specialize_available(), reason="no libnvrtc >= 12.8 in this container")


def _handle(prog):
    return DeviceProgramHandle.create(prog.flat, prog.consts, prog.arg_pool, prog.param_count, prog.workspace_size,
                                      prog.result_regs)


@needs_nvrtc
@pytest.mark.parametrize("name", ["single_expr", "clifford", "aizawa", "double_pendulum"])
def test_named_programs_compile_for_sm_100a(name):
    w = workloads.get(name)
    h = _handle(w.program)
    before = h.specialized_info()
    assert before.compiled == 0 and before.cubin_bytes == 0
    info = h.specialize()
    assert info.compiled == 1 and info.cubin_bytes > 1000 and info.compile_ms > 0
    assert info.points_per_thread in (1, 2, 4) and info.block_threads % 32 == 0
    assert info.registers == -1, "no device here: the kernel cannot have been loaded"
    assert info.can_iterate == (1 if w.iterated else 0) or not w.iterated
    again = h.specialize()  # idempotent: compiled once
    assert again.compile_ms == info.compile_ms


@needs_nvrtc
def test_generated_text_is_the_device_program_statement_by_statement():
    # C1: x^2 + sin(x)*exp(-x)  ->  Square, Sin, ExpNeg, MulAdd; four points per thread
    w = workloads.get("single_expr")
    h = _handle(w.program)
    info = h.specialize()
    full_src = h.specialized_source()
    # two entry points: saf_spec, and saf_spec_pf (single passes: next-tile prefetch) with the same statements
    assert info.has_prefetch_entry == 1 and full_src.count("__global__") == 2
    src = full_src[:full_src.index("saf_spec_pf")]
    assert "n0_0" in full_src[full_src.index("saf_spec_pf"):] and "n0_0" not in src
    n_instr = h.info().n_dev_instructions
    # the body is written twice: the speculative copy (fast paths only, sp::*_f) and the careful replay (sp::*_v)
    assert len(re.findall(r"// \[\d+\]", src)) == 2 * n_instr
    P = info.points_per_thread
    assert src.count("__dmul_rn(") >= 2 * P  # the square, once per point and copy
    assert src.count("sp::sin_v(") == 1 and src.count("sp::exp_v(") == 1 and src.count("__fma_rn(") == 2 * P
    assert src.count("sp::sin_f(") == 1 and src.count("sp::exp_f(") == 1
    assert "if (slow_t_ >= 0x41300000u || slow_e_ >= 0x4085e000u) {" in src
    assert "lm::" not in src and "for (int p" not in src, "statements are written out per point: no arrays, no loops"
    # constants travel as bit patterns
    for k in re.findall(r"const double k\d+ = (\S+);", src):
        assert k.startswith("__longlong_as_double(0x")
    assert f"#define SPEC_P {P}" in src and "#include \"spec_prelude.cuh\"" in src
    pend = _handle(workloads.get("double_pendulum").program)  # FP64 time only: no prefetch entry
    assert pend.specialize().has_prefetch_entry == 0


@needs_nvrtc
def test_constant_bits_survive_including_negative_zero_and_nan():
    prog = compile_exprs([parse("x + 0.1"), parse("x * 3.0")], ["x"])
    h = _handle(prog)
    h.specialize()
    src = h.specialized_source()
    assert f"0x{np.float64(0.1).view(np.uint64):016x}" in src


@needs_nvrtc
def test_iteration_feedback_is_emitted_only_when_results_pair_with_parameters():
    it = _handle(workloads.get("clifford").program)
    assert it.specialize().can_iterate == 1
    assert re.search(r"if \(\+\+it >= d\.iters\) break;\n\s+s\d+_0 = ", it.specialized_source())
    one = _handle(compile_exprs([parse("x*y")], ["x", "y"]))  # one result, two parameters
    assert one.specialize().can_iterate == 0
    assert not re.search(r"if \(\+\+it >= d\.iters\) break;\n\s+s\d+_0 = ", one.specialized_source())


@needs_nvrtc
def test_speculation_only_where_leaves_are_inlined_and_exist(monkeypatch):
    a = _handle(workloads.get("aizawa").program)  # no transcendental at all: one copy of the body
    a.specialize()
    assert "slow_t_" not in a.specialized_source()
    monkeypatch.setenv("SAF_SPEC_SPECULATE", "0")  # tuning knob
    c = _handle(workloads.get("clifford").program)
    c.specialize()
    assert "slow_t_" not in c.specialized_source() and "sp::sin_v(" in c.specialized_source()


@needs_nvrtc
def test_many_transcendentals_are_called_not_inlined():
    few = _handle(workloads.get("clifford").program)
    few.specialize()
    assert "#define SPEC_OUTLINE" not in few.specialized_source()
    terms = " + ".join(f"sin({1 + 0.1 * k}*x)*exp(-{0.2 + 0.01 * k}*x)" for k in range(20))
    many = _handle(compile_exprs([parse(terms)], ["x"]))
    many.specialize()
    assert "#define SPEC_OUTLINE 1" in many.specialized_source() and "slow_t_" not in many.specialized_source()


def test_programs_above_the_limit_are_refused_and_stay_usable(monkeypatch):
    monkeypatch.setenv("SAF_SPEC_MAX_INSTRUCTIONS", "10")  # the shipped limit is 20000 device instructions
    prog = workloads.get("aizawa").program
    h = _handle(prog)
    assert h.info().n_dev_instructions > 10
    with pytest.raises(SafError) as ei:
        h.specialize()
    assert ei.value.code == 6 and "interpreted" in ei.value.message
    assert h.specialized_info().compiled == 0
    assert h.info().n_dev_instructions > 10  # the handle is intact


# ---- the generated text, EXECUTED on the CPU (tests/spec_host: g++ + a CUDA-vocabulary shim; test infrastructure) ----------
from oracle import OracleProgram  # noqa: E402
from spec_host import host_kernel  # noqa: E402
from util import assert_close, bits_equal, random_program  # noqa: E402


def _host(prog, monkeypatch=None, **env):
    if monkeypatch is not None:
        for k, v in env.items():
            monkeypatch.setenv(k, str(v))
    h = _handle(prog)
    h.specialize()
    return host_kernel(h.specialized_source())


@needs_nvrtc
@pytest.mark.parametrize("entry", [0, 1])
@pytest.mark.parametrize("n", [1, 511, 512, 1537, 5000])
def test_generated_aizawa_kernel_is_bit_exact_against_the_oracle_on_the_cpu(n, entry):
    # arithmetic only: every statement is an IEEE operation -> the same bits as the reference interpreter restated
    w = workloads.get("aizawa")
    k = _host(w.program)
    assert k.points_per_thread == 4 and k.has_prefetch_entry
    cols = w.make_inputs(n, 9)
    ref = OracleProgram.from_program(w.program).eval_batch(cols, n)
    got = k.run(cols, n, 3, grid=3, entry=entry)  # three blocks: several tiles per block, a ragged last tile
    for g, r in zip(got, ref):
        assert bits_equal(g, r)


@needs_nvrtc
def test_generated_kernel_iterates_in_place_like_the_oracle():
    w = workloads.get("aizawa")
    k = _host(w.program)
    n = 3000
    cols = w.make_inputs(n, 2)
    ref = OracleProgram.from_program(w.program).iterate(cols, 7)
    st = [c.copy() for c in cols]
    k.run(st, n, 3, iters=7, grid=2, outs=st)  # outputs alias the columns: saf_iterate_device's call
    for g, r in zip(st, ref):
        assert bits_equal(g, r)


@needs_nvrtc
@pytest.mark.parametrize("name,p", [("single_expr", 4), ("clifford", 2), ("double_pendulum", 2), ("double_pendulum", 1)])
def test_generated_kernels_with_transcendentals_match_the_oracle(name, p, monkeypatch):
    w = workloads.get(name)
    k = _host(w.program, monkeypatch, SAF_SPEC_P=p)
    assert k.points_per_thread == p
    n = 2049
    cols = w.make_inputs(n, 4)
    n_out = len(w.program.result_regs)
    ref = OracleProgram.from_program(w.program).eval_batch(cols, n)
    for entry in ([0, 1] if k.has_prefetch_entry else [0]):
        got = k.run(cols, n, n_out, grid=2, entry=entry)
        for j, (g, r) in enumerate(zip(got, ref)):
            assert_close(g, r, 1e-12, 1e-13, what=f"{name} output {j} entry {entry}")


@needs_nvrtc
def test_speculative_body_replays_the_iteration_on_the_cpu():
    # arguments outside the fast range of the lean kernels: the fast copy of the body flags them, the careful copy
    # recomputes the whole iteration from the saved parameters (spec.cpp `speculate`)
    w = workloads.get("clifford")
    k = _host(w.program)
    n = 1024
    cols = w.make_inputs(n, 5)
    cols[0][::7] = 3.0e7
    cols[1][3::11] = -8.0e9
    cols[0][5] = np.inf
    cols[1][6] = np.nan
    o = OracleProgram.from_program(w.program)
    ref = o.eval_batch(cols, n)
    got = k.run(cols, n, 2, grid=1)
    for g, r in zip(got, ref):
        assert_close(g, r, 1e-12, 1e-13, what="clifford with slow-path arguments")
    ref3 = o.iterate(cols, 3)
    st = [c.copy() for c in cols]
    k.run(st, n, 2, iters=3, grid=1, outs=st)
    for g, r in zip(st, ref3):
        assert_close(g, r, 1e-11, 1e-13, what="clifford x3 with slow-path start")


@needs_nvrtc
def test_generated_kernel_column_rules_missing_short_unaligned():
    prog = compile_exprs([parse("a*b + c")], ["a", "b", "c"])
    k = _host(prog)
    rng = np.random.default_rng(3)
    n = 1300
    a, b, c = rng.normal(size=n), rng.normal(size=700), rng.normal(size=1)
    o = OracleProgram.from_program(prog)
    assert bits_equal(k.run([a, b, c], n, 1)[0], o.eval_batch([a, b, c], n)[0])          # short columns repeat their last value
    assert bits_equal(k.run([a, a, None], n, 1)[0], o.eval_batch([a, a], n)[0])          # missing column reads 0.0
    buf = np.empty(n + 1)
    buf[1:] = a
    out = np.empty(n + 1)
    got = k.run([buf[1:], b, c], n, 1, outs=[out[1:]])[0]                                  # 8-byte aligned only: scalar accesses
    assert bits_equal(got, o.eval_batch([a, b, c], n)[0])


@needs_nvrtc
@pytest.mark.parametrize("seed", range(8))
def test_generated_kernels_of_random_programs_match_the_oracle(seed):
    rng = np.random.default_rng(7000 + seed)
    prog = random_program(rng, n_params=int(rng.integers(1, 5)), n_instr=int(rng.integers(1, 40)), n_results=int(rng.integers(1, 4)))
    n = 300
    cols = [rng.uniform(-3, 3, n) for _ in range(prog.param_count)]
    ref = OracleProgram.from_program(prog).eval_batch(cols, n)
    got = _host(prog).run(cols, n, len(prog.result_regs), grid=2)
    for g, r in zip(got, ref):
        assert_close(g, r, rtol=1e-6, atol=1e-10, what=f"seed {seed}")  # the reference's own fuzz comparator


@needs_nvrtc
@pytest.mark.parametrize("small", [False, True])
def test_generated_long_program_kernel_matches_the_oracle_on_the_cpu(small, monkeypatch):
    # the terrain gradient takes the long-program path of the generator: leaves called (SPEC_OUTLINE), the two SinCos of a
    # term in one call (sincos_v2), block barriers between stretches of statements (no-ops here: threads run in turn)
    if small:
        monkeypatch.setenv("SAF_SPEC_P", "1")
        monkeypatch.setenv("SAF_SPEC_BLOCK", "128")
    w = workloads.get("terrain_gradient")
    h = _handle(w.program)
    info = h.specialize()
    src = h.specialized_source()
    assert "#define SPEC_OUTLINE 1" in src and src.count("sp::sincos_v2(") == 112 and "__syncthreads();" in src
    assert info.points_per_thread == (1 if small else 2) and info.has_prefetch_entry == 0
    k = host_kernel(src)
    n = 1500
    cols = w.make_inputs(n, 3)
    ref = OracleProgram.from_program(w.program).eval_batch(cols, n)
    got = k.run(cols, n, 2, grid=2)
    for j, (g, r) in enumerate(zip(got, ref)):
        assert_close(g, r, 1e-12, 1e-13, what=f"terrain gradient output {j}")


@needs_nvrtc
def test_rolled_kernel_of_a_long_sum_returns_the_bits_of_the_straight_line_kernel(monkeypatch):
    # csrc/reroll.inc (opt-in, SAF_SPEC_REROLL=1): the 110-term gradient as a loop over its terms, constants in a table
    w = workloads.get("terrain_gradient")
    straight = _handle(w.program)
    straight.specialize()
    monkeypatch.setenv("SAF_SPEC_REROLL", "1")
    rolled = _handle(w.program)
    info = rolled.specialize()
    src = rolled.specialized_source()
    assert "saf_rr_kt" in src and "for (int g_ = 0; g_ < 222; ++g_)" in src and "#define SPEC_OUTLINE" not in src
    assert info.block_threads == 128 and len(src) < len(straight.specialized_source()) // 2
    n = 1500
    cols = w.make_inputs(n, 3)
    a = host_kernel(straight.specialized_source()).run(cols, n, 2, grid=2)
    b = host_kernel(src).run(cols, n, 2, grid=2)
    for x, y in zip(a, b):
        assert bits_equal(x, y)
    # a program without long running sums is left alone
    c = _handle(workloads.get("double_pendulum").program)
    c.specialize()
    assert "saf_rr_kt" not in c.specialized_source()


def _random_expression(rng, depth=0):
    leaves = ["x", "y", "z", "0.5", "1.25", "-2.0", "3.0"]
    if depth >= 4 or rng.random() < 0.2:
        return leaves[int(rng.integers(len(leaves)))]
    kind = rng.random()
    a = _random_expression(rng, depth + 1)
    if kind < 0.45:
        b = _random_expression(rng, depth + 1)
        return f"({a} {'+-*/'[int(rng.integers(4))]} {b})"
    if kind < 0.85:
        fn = ["sin", "cos", "exp", "tanh", "sinh", "atan", "erf", "sqrt(abs", "log1p(abs", "cosh"][int(rng.integers(10))]
        return f"{fn}({a}))" if "(" in fn else f"{fn}({a})"
    return f"({a})^{int(rng.integers(2, 5))}"


@needs_nvrtc
@pytest.mark.parametrize("seed", range(12))
def test_generated_kernels_of_random_expressions_match_the_oracle(seed):
    # expression strings -> C++ compiler (csrc/compile.cpp) -> lowering -> generated kernel, executed on the CPU, against
    # the oracle running the same reference bytecode
    rng = np.random.default_rng(9000 + seed)
    exprs = [_random_expression(rng) for _ in range(int(rng.integers(1, 4)))]
    h = DeviceProgramHandle.compile(exprs, ["x", "y", "z"])
    flat, consts, pool, param_count, workspace, regs = h.reference_payload(0)
    from symbanafis_b200.isa import Program
    prog = Program(flat=list(flat), consts=list(consts), arg_pool=list(pool), param_count=param_count, workspace_size=workspace,
                   result_regs=list(regs))
    h.specialize()
    n = 257
    cols = [rng.uniform(-1.5, 1.5, n) for _ in range(3)]
    ref = OracleProgram.from_program(prog).eval_batch(cols, n)
    got = host_kernel(h.specialized_source()).run(cols, n, len(exprs), grid=2)
    for e, g, r in zip(exprs, got, ref):
        assert_close(g, r, rtol=1e-9, atol=1e-12, what=e)


@needs_nvrtc
def test_disk_cache_of_compiled_kernels_is_opt_in_and_keyed_by_content(tmp_path, monkeypatch):
    import os
    w = workloads.get("aizawa")
    a = _handle(w.program)
    a.specialize()
    assert not list(tmp_path.iterdir())                      # no cache directory configured: nothing is written
    monkeypatch.setenv("SAF_SPEC_CACHE_DIR", str(tmp_path))
    b = _handle(w.program)
    ib = b.specialize()
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].startswith("saf_spec_sm100a_nvrtc") and files[0].endswith(".cubin")
    c = _handle(w.program)
    ic = c.specialize()                                      # served from the file: no NVRTC
    assert ic.cubin_bytes == ib.cubin_bytes and ic.compile_ms < ib.compile_ms / 3
    d = _handle(workloads.get("clifford").program)           # another program: another entry
    d.specialize()
    assert len(os.listdir(tmp_path)) == 2
    # a damaged entry is ignored and replaced
    with open(tmp_path / files[0], "wb") as f:
        f.write(b"garbage")
    e = _handle(w.program)
    assert e.specialize().cubin_bytes == ib.cubin_bytes
    assert os.path.getsize(tmp_path / files[0]) == ib.cubin_bytes


@needs_nvrtc
def test_rolling_other_long_sums_or_declining_them(monkeypatch):
    from symbanafis_b200.expr import diff
    rng = np.random.default_rng(5)
    g = " + ".join(f"{rng.uniform(0.2, 1):.6f}*exp(-((x - {rng.uniform(-2, 2):.5f})^2 + (y - {rng.uniform(-2, 2):.5f})^2)/"
                   f"{rng.uniform(0.3, 2):.5f})" for _ in range(120))
    e = parse(g)
    mixture = compile_exprs([diff(e, "x"), diff(e, "y"), e], ["x", "y"])   # a Gaussian mixture, its gradient: three sums
    poly = " + ".join(f"{rng.uniform(-1, 1):.6f}*x^{k % 7 + 1}*y^{k % 5 + 1}/(1 + {rng.uniform(0.1, 2):.5f}*x^2)" for k in range(150))
    rational = compile_exprs([parse(poly)], ["x", "y"])                    # 35 different term shapes: not worth a loop
    n = 700
    cols = [rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)]
    for prog, expect_rolled in ((mixture, True), (rational, False)):
        monkeypatch.delenv("SAF_SPEC_REROLL", raising=False)
        straight = _handle(prog)
        straight.specialize()
        monkeypatch.setenv("SAF_SPEC_REROLL", "1")
        rolled = _handle(prog)
        rolled.specialize()
        assert ("saf_rr_kt" in rolled.specialized_source()) == expect_rolled
        a = host_kernel(straight.specialized_source()).run(cols, n, len(prog.result_regs), grid=2)
        b = host_kernel(rolled.specialized_source()).run(cols, n, len(prog.result_regs), grid=2)
        for x, y in zip(a, b):
            assert bits_equal(x, y)
        ref = OracleProgram.from_program(prog).eval_batch(cols, n)
        for x, r in zip(b, ref):
            assert_close(x, r, 1e-11, 1e-13, what="rolled vs oracle")
