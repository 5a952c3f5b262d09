#!/usr/bin/env python
"""Small-n replay of the GPU parity suite's shapes for compute-sanitizer (tools/gpu_sanitize.sh): every layout (one,
two, four points per thread, wide blocks through device-resident columns), iterated maps (double-buffered images and
the copy-back path), cold micro-ops (builtins), multi-window programs, ragged tails, broadcast and missing columns.
Results are checked against the oracle so that a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import OracleProgram  # noqa: E402
from symbanafis_b200 import workloads  # noqa: E402
from symbanafis_b200._lib import DeviceProgramHandle  # noqa: E402
from util import close_enough, random_program  # noqa: E402


def handle(p):
    return DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)


def check(got, ref, what):
    for g, r in zip(got, ref):
        assert close_enough(g, r, 1e-9, 1e-12).all(), what


def main():
    n_cases = 0
    for name in workloads.WORKLOADS:
        w = workloads.get(name)
        h = handle(w.program)
        op = OracleProgram.from_program(w.program)
        for n in (1, 257, 3001):
            cols = w.make_inputs(n, 3)
            outs = [np.empty(n) for _ in w.program.result_regs]
            h.eval_host(cols, outs, n)
            check(outs, op.run_chunked(cols, n), f"{name} n={n}")
            n_cases += 1
        if w.iterated:
            cols = w.make_inputs(1500, 4)
            st = [c.copy() for c in cols]
            h.iterate_host(st, 3)
            check(st, op.iterate(cols, 3), f"{name} iterate")
            n_cases += 1
    rng = np.random.default_rng(99)
    for seed in range(12):
        prog = random_program(rng, n_params=int(rng.integers(1, 5)), n_instr=int(rng.integers(5, 40)))
        n = int(rng.integers(1, 2000))
        cols = [rng.uniform(-3, 3, n) for _ in range(prog.param_count)]
        outs = [np.empty(n) for _ in prog.result_regs]
        handle(prog).eval_host(cols, outs, n)
        check(outs, OracleProgram.from_program(prog).run_chunked(cols, n), f"random {seed}")
        n_cases += 1
    # broadcast / missing columns and an unaligned column (one point per thread path)
    w = workloads.get("aizawa")
    h = handle(w.program)
    n = 1001
    cols = w.make_inputs(n + 1, 5)
    short = [cols[0][1:], cols[1][:7], cols[2][:1]]
    outs = [np.empty(n) for _ in w.program.result_regs]
    h.eval_host(short, outs, n, broadcast=True)
    check(outs, OracleProgram.from_program(w.program).eval_batch(short, n), "broadcast")
    n_cases += 1
    # device-resident columns: one launch with enough points for the wide one-block-per-SM layouts
    import torch
    w = workloads.get("terrain_gradient")
    h = handle(w.program)
    n = int(os.environ.get("SAF_SANITIZE_WIDE_POINTS", "620000"))
    cols = w.make_inputs(n, 6)
    dev = torch.device("cuda", 0)
    tc = [torch.from_numpy(c).to(dev) for c in cols]
    to = [torch.empty(n, dtype=torch.float64, device=dev) for _ in w.program.result_regs]
    h.eval_device([t.data_ptr() for t in tc], [n] * len(tc), [t.data_ptr() for t in to], n)
    torch.cuda.synchronize()
    idx = np.arange(0, n, 997)
    ref = OracleProgram.from_program(w.program).eval_batch([c[idx] for c in cols], idx.size)
    check([t.cpu().numpy()[idx] for t in to], ref, "wide layout")
    n_cases += 1
    print(f"sanitize_cases: {n_cases} cases ok")


if __name__ == "__main__":
    main()
