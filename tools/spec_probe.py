"""Interpreter vs specialised kernel on the five configs: bits and device time (development probe, run on the GPU box)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symbanafis_b200 import _lib, workloads  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["single_expr", "clifford", "aizawa", "double_pendulum", "terrain_gradient"]
SIZES = {"single_expr": (int(os.environ.get("SINGLE_N", 1_000_000)), 1), "clifford": (1_000_000, 1000), "aizawa": (500_000, 600), "double_pendulum": (int(os.environ.get("PEND_N", 50_000)), 600),
         "terrain_gradient": (int(os.environ.get("TERRAIN_N", 20_000_000)), 1)}
peak = _lib.measure_fp64_peak()


def handle(w):
    p = w.program
    return _lib.DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)


def run(h, w, cols, n, iters, steps=5):
    cols0 = [torch.from_numpy(c).to(dev) for c in cols]
    n_out = len(w.program.result_regs)
    state = [c.clone() for c in cols0]
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(n_out)]
    s = torch.cuda.current_stream().cuda_stream

    def launch():
        if w.iterated:
            for t, c in zip(state, cols0):
                t.copy_(c)
            h.iterate_device([t.data_ptr() for t in state], n, iters, stream=s)
        else:
            h.eval_device([c.data_ptr() for c in cols0], [n] * len(cols0), [o.data_ptr() for o in outs], n, stream=s)

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.fill_(1.0)
        if w.iterated:
            for t, c in zip(state, cols0):
                t.copy_(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if w.iterated:
            h.iterate_device([t.data_ptr() for t in state], n, iters, stream=s)
        else:
            h.eval_device([c.data_ptr() for c in cols0], [n] * len(cols0), [o.data_ptr() for o in outs], n, stream=s)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    res = [t.cpu().numpy() for t in (state if w.iterated else outs)]
    return min(ms), sum(ms) / len(ms), res


for name in names:
    w = workloads.get(name)
    n, iters = SIZES[name]
    cols = w.make_inputs(n, 7)
    hi = handle(w)
    info = hi.info()
    os.environ.pop("SAF_SPECIALIZE", None)
    t_i = run(hi, w, cols, n, iters)
    hs = handle(w)
    t0 = time.time()
    si = hs.specialize()
    t_s = run(hs, w, cols, n, iters)
    si = hs.specialized_info()
    same = all(np.array_equal(a.view(np.uint64), b.view(np.uint64)) for a, b in zip(t_i[2], t_s[2]))
    nan_same = all(np.array_equal(np.isnan(a), np.isnan(b)) for a, b in zip(t_i[2], t_s[2]))
    flops = info.fp64_slots * n * iters
    print(json.dumps({"workload": name, "n": n, "iters": iters, "interp_ms": round(t_i[1], 4), "spec_ms": round(t_s[1], 4),
                      "spec_min_ms": round(t_s[0], 4), "speedup": round(t_i[1] / t_s[1], 2), "bits_equal": bool(same), "nan_pattern_equal": bool(nan_same),
                      "interp_frac_fp64": round(flops / (t_i[1] * 1e-3) / peak, 3), "spec_frac_fp64": round(flops / (t_s[1] * 1e-3) / peak, 3),
                      "hbm_GBs_spec": round(info.bytes_per_point * n / (t_s[1] * 1e-3) / 1e9, 1),
                      "P": si.points_per_thread, "block": si.block_threads, "regs": si.registers, "pf_regs": si.prefetch_registers, "small": [si.small_compiled, si.small_points_per_thread, si.small_registers], "local": si.local_bytes, "resident": si.resident_blocks,
                      "compile_ms": round(si.compile_ms, 1), "env": {k: v for k, v in os.environ.items() if k.startswith("SAF_SPEC")}}), flush=True)
