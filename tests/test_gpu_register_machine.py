"""The register machine (csrc/ruop.h, kernel_r.cuh, pack_r.cpp) is opt-in (SAF_MACHINE=r, read once per process):
the GPU parity suite is replayed on it in a child process, and its results must equal the shared-memory
machine's bit for bit (both run the same correctly rounded operations in the same order)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from symbanafis_b200 import workloads
from symbanafis_b200._lib import DeviceProgramHandle
out = {{}}
for name, n, iters in (("aizawa", 70001, 1), ("aizawa", 70001, 7), ("double_pendulum", 5003, 3), ("clifford", 40002, 5),
                       ("single_expr", 100001, 1), ("terrain_gradient", 3001, 1)):
    w = workloads.get(name); p = w.program
    h = DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)
    cols = w.make_inputs(n, 5)
    if iters == 1:
        res = [np.empty(n) for _ in p.result_regs]
        h.eval_host(cols, res, n)
    else:
        res = [c.copy() for c in cols]
        h.iterate_host(res, iters)
    for k, r in enumerate(res):
        out[f"{{name}}_{{iters}}_{{k}}"] = r
np.savez(sys.argv[1], **out)
"""


def _run_child(machine: str, path: str, **extra_env) -> None:
    env = dict(os.environ, SAF_MACHINE=machine, **extra_env)
    subprocess.run([sys.executable, "-c", _CHILD.format(root=ROOT), path], check=True, env=env, timeout=600)


_CHILD_WIDE = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from symbanafis_b200 import workloads
from symbanafis_b200._lib import DeviceProgramHandle
w = workloads.get("terrain_gradient"); p = w.program
h = DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)
import torch
n = 700_001  # one launch with enough points for the wide one-block-per-SM layouts (P = 4 and P = 2), ragged last tile
cols = w.make_inputs(n, 8)
dev = torch.device("cuda", 0)
tc = [torch.from_numpy(c).to(dev) for c in cols]
to = [torch.empty(n, dtype=torch.float64, device=dev) for _ in p.result_regs]
h.eval_device([t.data_ptr() for t in tc], [n] * len(tc), [t.data_ptr() for t in to], n)
torch.cuda.synchronize()
np.savez(sys.argv[1], **{{f"r{{k}}": t.cpu().numpy() for k, t in enumerate(to)}})
"""


@pytest.mark.gpu
def test_layout_and_fusion_knobs_do_not_change_results(tmp_path):
    # wide blocks (kernel.cu WIDE), the fused-chain peephole (pack.cpp) and the number of points per thread only
    # change HOW the same correctly rounded operations are dispatched
    outs = {}
    for tag, env in (("default", {}), ("narrow", {"SAF_WIDE": "0"}), ("unfused", {"SAF_FUSE": "0"}),
                     ("p1", {"SAF_POINTS_PER_THREAD": "1"}), ("p2wide", {"SAF_POINTS_PER_THREAD": "2"}),
                     ("p2narrow", {"SAF_POINTS_PER_THREAD": "2", "SAF_WIDE": "0"}), ("sched_a", {"SAF_SCHED": "n"})):
        path = str(tmp_path / f"{tag}.npz")
        subprocess.run([sys.executable, "-c", _CHILD_WIDE.format(root=ROOT), path], check=True,
                       env=dict(os.environ, **env), timeout=900)
        outs[tag] = np.load(path)
    for tag in ("narrow", "unfused", "p1", "p2wide", "p2narrow", "sched_a"):
        for k in outs["default"].files:
            assert np.array_equal(outs["default"][k].view(np.uint64), outs[tag][k].view(np.uint64)), f"{tag}: {k} differs"


@pytest.mark.gpu
def test_register_machine_equals_shared_memory_machine_bit_for_bit(tmp_path):
    a, b = str(tmp_path / "s.npz"), str(tmp_path / "r.npz")
    _run_child("s", a)
    _run_child("r", b)
    sa, sb = np.load(a), np.load(b)
    assert sorted(sa.files) == sorted(sb.files) and sa.files
    for k in sa.files:
        assert np.array_equal(sa[k].view(np.uint64), sb[k].view(np.uint64)), f"{k}: register machine differs"


@pytest.mark.gpu
def test_parity_suite_on_the_register_machine():
    env = dict(os.environ, SAF_MACHINE="r")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x",
                        "-q", "-p", "no:cacheprovider"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
