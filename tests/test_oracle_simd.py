"""The f64x4 restatement of the reference's SIMD path (oracle/saf_oracle_simd.c, simd.rs:53-114) against the scalar
restatement: same bits on the five named programs, on random programs over the whole ISA, on ragged sizes around the
4-lane groups and the 256-point chunks, and for iterated maps."""
import numpy as np
import pytest

import oracle
from oracle import OracleProgram
from symbanafis_b200 import workloads
from util import bits_equal, random_program


@pytest.mark.parametrize("name", list(workloads.WORKLOADS))
@pytest.mark.parametrize("n", [256, 257, 259, 1023, 4100])
def test_simd_body_equals_scalar_body_on_named_programs(name, n):
    w = workloads.get(name)
    op = OracleProgram.from_program(w.program)
    cols = w.make_inputs(n, 11)
    a = op.run_chunked(cols, n, 3)
    b = op.run_chunked_simd(cols, n, 3)
    for x, y in zip(a, b):
        assert bits_equal(x, y)


@pytest.mark.parametrize("seed", range(30))
def test_simd_body_equals_scalar_body_on_random_programs(seed):
    rng = np.random.default_rng(7000 + seed)
    prog = random_program(rng, n_params=int(rng.integers(1, 5)), n_instr=int(rng.integers(1, 40)))
    n = int(rng.integers(256, 1500))
    cols = [rng.uniform(-3, 3, n) for _ in range(prog.param_count)]
    op = OracleProgram.from_program(prog)
    for x, y in zip(op.run_chunked(cols, n, 2), op.run_chunked_simd(cols, n, 2)):
        assert bits_equal(x, y)


@pytest.mark.parametrize("name", ["clifford", "aizawa", "double_pendulum"])
def test_simd_iteration_equals_scalar_iteration(name):
    w = workloads.get(name)
    op = OracleProgram.from_program(w.program)
    cols = w.make_inputs(1031, 5)
    for x, y in zip(op.iterate(cols, 7, 2), op.iterate(cols, 7, 2, simd=True)):
        assert bits_equal(x, y)


def test_small_batches_take_the_scalar_loop_and_errors_are_the_same():
    w = workloads.get("single_expr")
    op = OracleProgram.from_program(w.program)
    x = np.linspace(-2, 2, 100)
    assert bits_equal(op.run_chunked_simd([x], 100)[0], op.run_chunked([x], 100)[0])
    with pytest.raises(oracle.OracleError):
        op.run_chunked_simd([x], 300)  # column length mismatch (batch.rs:48-50)
    assert isinstance(oracle.has_simd(), bool)
