"""Host half of the product on CPU: the lowering (value numbering, chain expansion, scheduling,
accumulator forwarding, slot allocation) must preserve results BIT FOR BIT.

The device program exported by the C ABI is run by tests/dev_emulator.py (kernel state model, oracle
leaf math) and compared with the oracle executing the original reference bytecode."""
import numpy as np
import pytest

from dev_emulator import R_LAYOUT, run_device_program, run_packed_program, run_packed_r_program
from oracle import OracleProgram
from symbanafis_b200 import workloads
from symbanafis_b200._lib import DeviceProgramHandle, SafError
from util import bits_equal, random_program


def _handle(prog):
    return DeviceProgramHandle.create(prog.flat, prog.consts, prog.arg_pool, prog.param_count,
                                      prog.workspace_size, prog.result_regs)


def _check(prog, cols, n, iters=1):
    h = _handle(prog)
    code, consts, pslot, rfield = h.export()
    o = OracleProgram.from_program(prog)
    ref = o.eval_batch(cols, n) if iters == 1 else o.iterate(cols, iters)
    got = run_device_program(code, consts, pslot, rfield, cols, n, iters)
    for k, (r, g) in enumerate(zip(ref, got)):
        assert bits_equal(g, r), f"result {k} differs"
    # the packed micro-instruction image the kernel executes, for every register-file layout
    for ppt in (1, 2, 4, 8, 2000 + 320, 4000 + 512):  # 1000 * P + T: P points per thread, one wide block of T threads per SM
        got = run_packed_program(h.export_packed(ppt), cols, n, iters)
        for k, (r, g) in enumerate(zip(ref, got)):
            assert bits_equal(g, r), f"packed image (P={ppt}): result {k} differs"
    # the register machine's image (pack_r.cpp: register allocation, spills, state registers of iterated maps);
    # it loops in-kernel only when results pair with parameters
    if iters == 1 or (len(rfield) == prog.param_count and len(rfield) <= 5):
        got = run_packed_r_program(h.export_packed(R_LAYOUT), cols, n, iters)
        for k, (r, g) in enumerate(zip(ref, got)):
            assert bits_equal(g, r), f"register-machine image: result {k} differs"
    return h


@pytest.mark.parametrize("name", list(workloads.WORKLOADS))
def test_workload_lowering_is_bit_exact(name):
    w = workloads.get(name)
    cols = w.make_inputs(96, 11)
    h = _check(w.program, cols, 96)
    info = h.info()
    assert info.n_results == len(w.program.result_regs)
    assert info.n_slots <= 64, "scheduler must keep the on-chip register file small"


@pytest.mark.parametrize("name", ["clifford", "aizawa", "double_pendulum"])
def test_iterated_lowering_is_bit_exact(name):
    w = workloads.get(name)
    cols = w.make_inputs(40, 5)
    _check(w.program, cols, 40, iters=4)


@pytest.mark.parametrize("seed", range(60))
def test_random_programs(seed):
    rng = np.random.default_rng(seed)
    prog = random_program(rng, n_params=int(rng.integers(1, 5)), n_instr=int(rng.integers(1, 40)),
                          n_results=int(rng.integers(1, 4)))
    n = 33
    cols = [rng.uniform(-3, 3, n) for _ in range(prog.param_count)]
    _check(prog, cols, n)


def test_result_is_parameter_or_constant():
    from symbanafis_b200.isa import Program, assemble
    # program with no instructions: results are a parameter and a constant
    prog = Program(flat=assemble([]), consts=[2.5], arg_pool=[], param_count=2, workspace_size=3, result_regs=[1, 2, 1])
    rng = np.random.default_rng(0)
    cols = [rng.normal(size=17), rng.normal(size=17)]
    _check(prog, cols, 17)
    # swap map (x, y) -> (y, x) iterated: result slots must not alias parameter slots
    swap = Program(flat=assemble([]), consts=[], arg_pool=[], param_count=2, workspace_size=2, result_regs=[1, 0])
    _check(swap, cols, 17, iters=3)


def test_fuse_matches_separate_programs():
    w = workloads.get("aizawa")
    hs = [_handle(p) for p in w.parts]
    fused = DeviceProgramHandle.fuse(hs)
    info = fused.info()
    assert info.n_results == 3
    assert info.n_dev_instructions < sum(h.info().n_dev_instructions for h in hs)  # (z - 0.7) is shared
    cols = w.make_inputs(50, 3)
    code, consts, pslot, rfield = fused.export()
    got = run_device_program(code, consts, pslot, rfield, cols, 50)
    got4 = run_packed_program(fused.export_packed(4), cols, 50)
    for p, g, g4 in zip(w.parts, got, got4):
        ref = OracleProgram.from_program(p).eval_batch(cols, 50)[0]
        assert bits_equal(g, ref) and bits_equal(g4, ref)


def test_validator_rejects_malformed_programs():
    from symbanafis_b200.isa import assemble
    bad = [
        (assemble([("Add", 5, 0, 1)]), 2, 3, [2]),               # destination out of bounds
        (assemble([("Add", 2, 0, 3)]), 2, 4, [2]),               # read before definition
        (np.asarray([4, 2, 0, 1], dtype=np.uint32), 2, 3, [2]),  # no End sentinel
        (np.asarray([99, 0], dtype=np.uint32), 2, 3, [0]),       # unknown opcode
        (assemble([("AddN", 2, 0, 3)]), 2, 3, [2]),              # arg pool out of range
        (assemble([("Builtin1", 2, 200, 0)]), 2, 3, [2]),        # unknown FnOp
        (assemble([("Add", 2, 0, 1)]), 2, 3, [7]),               # result register out of bounds
    ]
    for flat, pc, ws, res in bad:
        with pytest.raises(SafError) as ei:
            DeviceProgramHandle.create(flat, [], [], pc, ws, res)
        assert ei.value.code == 2
        with pytest.raises(Exception):
            OracleProgram(flat, [], [], pc, ws, res)


@pytest.mark.parametrize("seed", range(100, 112))
def test_register_machine_random_programs_with_feedback(seed):
    # pack_r.cpp: register pressure (up to 90 instructions over 8 registers), cold micro-ops with every live register
    # stored, state registers + parallel move at the end of the body when results pair with parameters
    rng = np.random.default_rng(seed)
    n_params = int(rng.integers(1, 7))
    n_results = n_params if seed % 2 else int(rng.integers(1, 5))
    prog = random_program(rng, n_params=n_params, n_instr=int(rng.integers(5, 90)), n_results=n_results,
                          builtins=bool(seed % 3))
    n = 40
    cols = [rng.uniform(-2, 2, n) for _ in range(n_params)]
    pk = _handle(prog).export_packed(R_LAYOUT)
    o = OracleProgram.from_program(prog)
    for iters in ((1, 2, 3) if (n_results == n_params and n_results <= 5) else (1,)):
        ref = o.eval_batch(cols, n) if iters == 1 else o.iterate([c.copy() for c in cols], iters)
        got = run_packed_r_program(pk, cols, n, iters)
        for k, (r, g) in enumerate(zip(ref, got)):
            assert bits_equal(g, r), f"register-machine image, {iters} iterations: result {k} differs"


def test_fused_chains_cut_the_dispatch_count():
    # pack.cpp peephole (uop.h U_FM / U_FQ / U_FS / U_FT): regression guard on the micro-op counts the round measured
    def micro_ops(name, ppt):
        pk = _handle(workloads.get(name).program).export_packed(ppt)
        n = (len(pk["image"]) - pk["code_off"] // 8) // 4
        return n // 2 if pk["double_buffered"] else n  # incl. the U_END of one copy

    assert micro_ops("clifford", 4) == 3 and micro_ops("clifford", 2) == 7  # sin/cos + add/fma fusion: P = 4 layout only
    assert micro_ops("terrain_gradient", 2) <= 1300  # 2666 device instructions
    assert micro_ops("double_pendulum", 1) <= 155    # 182
    assert micro_ops("aizawa", 4) <= 20              # 25
