#!/usr/bin/env python
"""bench.py -- f64 point-evals/s of the SymbAnaFis batch evaluator on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload =
BASELINE.json configs[1]: the Clifford attractor map, 1M particles x 1000 iterations per GPU, i.e. one
step = 1e9 point-evals per GPU (one point-eval = the whole 2-output program on one particle).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

N > 1 is launched by torchrun (one rank per GPU); points are sharded per rank, there is no collective
on the data path, the timing is the max over ranks.  Rank 0 prints ONE JSON line.

`value`  : device-timed (CUDA events on the launching stream), columns resident in HBM.
`e2e`    : the same step through the C-ABI host call (saf_iterate_host / saf_eval_host) with pinned
           HOST buffers: H2D of the inputs and D2H of the results inside the timed region.
`--impl reference`: the reference's CPU algorithm (oracle port, all host threads) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "f64 point-evals/sec"
UNIT = "point-evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="clifford")
    ap.add_argument("--points", type=int, default=0, help="points per GPU (0 = BASELINE.json size)")
    ap.add_argument("--iters", type=int, default=0, help="map iterations per step (0 = BASELINE.json value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def measured_traffic(workload_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of the interpreter launch from the committed ncu capture
    (profiles/r01_dram_traffic.json), or None when that workload has not been captured."""
    path = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f).get(workload_name)
        return None if t is None else t["dram_bytes_read"] + t["dram_bytes_write"]
    except (OSError, ValueError, KeyError):
        return None


def workload_setup(args):
    from symbanafis_b200 import workloads
    w = workloads.get(args.workload)
    n = args.points or w.n_points
    iters = args.iters or (w.iterations if w.iterated else 1)
    return w, n, iters


def cpu_baseline(w, n, iters, budget_s=12.0):
    """The oracle port (256-point chunks over all host cores) on a bounded sample of the same workload."""
    import oracle
    from oracle import OracleProgram
    op = OracleProgram.from_program(w.program)
    cores = oracle.online_cores()
    n_s = min(n, 1_000_000)
    cols = w.make_inputs(n_s, 1)
    # calibrate on a small run, then size the sample for ~budget_s of CPU work
    t0 = time.perf_counter()
    if w.iterated:
        op.iterate(cols, 2, cores)
    else:
        op.run_chunked(cols, n_s, cores)
    dt = time.perf_counter() - t0
    per_unit = dt / (n_s * (2 if w.iterated else 1))
    if w.iterated:
        it_s = int(max(2, min(iters, budget_s / max(per_unit * n_s, 1e-9))))
        t0 = time.perf_counter()
        op.iterate(cols, it_s, cores)
        dt = time.perf_counter() - t0
        units = n_s * it_s
        sample = f"{n_s} points x {it_s} iterations of {w.name}"
    else:
        reps = int(max(1, min(50, budget_s / max(per_unit * n_s, 1e-9))))
        t0 = time.perf_counter()
        for _ in range(reps):
            op.run_chunked(cols, n_s, cores)
        dt = time.perf_counter() - t0
        units = n_s * reps
        sample = f"{reps} passes over {n_s} points of {w.name}"
    return {"value": units / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "seconds": round(dt, 3)}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    w, n, iters = workload_setup(args)
    import oracle
    from oracle import OracleProgram
    op = OracleProgram.from_program(w.program)
    cores = oracle.online_cores()
    n_s = min(n, 1_000_000)
    cols = w.make_inputs(n_s, 1)
    # size one step for about 2 s of CPU work
    t0 = time.perf_counter()
    if w.iterated:
        op.iterate(cols, 1, cores)
    else:
        op.run_chunked(cols, n_s, cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    it_s = int(max(1, min(iters, 2.0 / dt))) if w.iterated else 1
    reps = 1 if w.iterated else int(max(1, min(20, 2.0 / dt)))

    def step():
        if w.iterated:
            op.iterate(cols, it_s, cores)
        else:
            for _ in range(reps):
                op.run_chunked(cols, n_s, cores)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    units = n_s * (it_s if w.iterated else reps)
    value = units * args.steps / dt
    sample = (f"{n_s} points x {it_s} iterations per step" if w.iterated else f"{reps} passes over {n_s} points per step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w.name}: {w.description}", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference interpreter (oracle/), all host cores; the Rust reference "
                "cannot be built in this image (no cargo/rustc)",
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from symbanafis_b200 import _lib

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the evaluator has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w, n, iters = workload_setup(args)
    p = w.program
    h = _lib.DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)
    info = h.info()
    n_state = p.param_count
    n_out = len(p.result_regs)

    # synthetic inputs: each rank owns its own contiguous shard of the global point set
    cols_np = w.make_inputs(n, 1 + rank)
    cols0 = [torch.from_numpy(c).to(dev) for c in cols_np]
    state = [c.clone() for c in cols0]
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(n_out)]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MiB > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def device_step():
        if w.iterated:
            for s, c in zip(state, cols0):
                s.copy_(c)
            h.iterate_device([s.data_ptr() for s in state], n, iters, stream=stream.cuda_stream)
        else:
            h.eval_device([c.data_ptr() for c in cols0], [n] * n_state, [o.data_ptr() for o in outs], n,
                          stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    fp64_peak = _lib.measure_fp64_peak()  # DFMA lane-instr/s, measured live on this GPU
    ideal = None
    if rank == 0 and w.name == "clifford":
        # the same map hard-coded around the same sin/cos kernels, no interpreter: what interpretation costs
        t_ideal = _lib.measure_clifford_probe(n, iters)
        ideal = {"kernel": "saf_clifford_probe_kernel (hard-coded map, state in registers, same launch shape)",
                 "ms_per_step": 1e3 * t_ideal, "value": n * iters / t_ideal, "unit": UNIT}

    for _ in range(max(3, args.warmup)):
        device_step()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.kernel_launch_count()
    kernel_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1.0)  # evict L2 between timed steps (not timed)
        if w.iterated:
            for s, c in zip(state, cols0):
                s.copy_(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if w.iterated:
            h.iterate_device([s.data_ptr() for s in state], n, iters, stream=stream.cuda_stream)
        else:
            h.eval_device([c.data_ptr() for c in cols0], [n] * n_state, [o.data_ptr() for o in outs], n,
                          stream=stream.cuda_stream)
        e1.record(stream)
        e1.synchronize()
        kernel_ms.append(e0.elapsed_time(e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.kernel_launch_count() - launches0
    step_ms = sum(kernel_ms) / len(kernel_ms)

    # ---- e2e: the host-buffer C-ABI call, pinned memory, copies inside the timed region ----
    pinned_in = [torch.from_numpy(c).pin_memory() for c in cols_np]
    pinned_state = [torch.empty_like(t).pin_memory() for t in pinned_in]
    pinned_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(n_out)]

    def host_step():
        if w.iterated:
            for s, c in zip(pinned_state, pinned_in):
                s.copy_(c)
            h.iterate_host([s.numpy() for s in pinned_state], iters)
        else:
            h.eval_host([c.numpy() for c in pinned_in], [o.numpy() for o in pinned_out], n)

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        host_step()
    barrier()
    e2e_s = []
    for _ in range(e2e_steps):
        if w.iterated:
            for s, c in zip(pinned_state, pinned_in):
                s.copy_(c)
        t0 = time.perf_counter()
        if w.iterated:
            h.iterate_host([s.numpy() for s in pinned_state], iters)
        else:
            h.eval_host([c.numpy() for c in pinned_in], [o.numpy() for o in pinned_out], n)
        e2e_s.append(time.perf_counter() - t0)
    barrier()
    e2e_ms = 1e3 * sum(e2e_s) / len(e2e_s)
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks ----
    from symbanafis_b200.sharding import max_over_ranks
    step_ms, e2e_ms, t_wall = max_over_ranks([step_ms, e2e_ms, t_wall], device=dev)

    units_per_step_gpu = n * iters
    units_per_step = units_per_step_gpu * world
    value = units_per_step / (step_ms * 1e-3)
    e2e_value = units_per_step / (e2e_ms * 1e-3)

    if rank == 0:
        hbm_gbs, peak_src = load_peaks()
        bytes_launch = info.bytes_per_point * n                # columns read + written once per launch
        slots_launch = info.fp64_slots * units_per_step_gpu    # FP64 issue slots per launch
        t_hbm = bytes_launch / (hbm_gbs * 1e9)
        t_fp64 = slots_launch / fp64_peak if fp64_peak > 0 else 0.0
        dur = step_ms * 1e-3
        if t_fp64 >= t_hbm:
            achieved = 2.0 * slots_launch / dur / 1e12
            peak = 2.0 * fp64_peak / 1e12
            roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak if peak else None,
                        "traffic": measured_traffic(w.name) if (n == w.n_points) else None,
                        "traffic_unit": "DRAM bytes per launch (ncu --set full capture, profiles/r01_dram_traffic.json)",
                        "peak_source": "DFMA-chain probe measured in this run (saf_measure_fp64_peak); "
                                       "FP64 is not in MEASURED_PEAKS.json",
                        "algorithmic": f"{info.fp64_slots:g} FP64 issue slots/point-eval (FMA = 2 FLOP) x "
                                       f"{units_per_step_gpu} point-evals/launch",
                        "hbm_leg": {"achieved_gbs": bytes_launch / dur / 1e9, "peak_gbs": hbm_gbs,
                                    "peak_source": peak_src}}
        else:
            achieved = bytes_launch / dur / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                        "frac": achieved / hbm_gbs, "traffic": measured_traffic(w.name) if (n == w.n_points) else None,
                        "peak_source": f"MEASURED_PEAKS.json ({peak_src})",
                        "algorithmic": f"{info.bytes_per_point:g} B/point x {n} points/launch",
                        "fp64_leg": {"achieved_tflops": 2.0 * slots_launch / dur / 1e12,
                                     "peak_tflops": 2.0 * fp64_peak / 1e12}}
        cb = None
        if not args.no_cpu_baseline:
            cb = cpu_baseline(w, n, iters)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{w.name}: {w.description}",
                "points_per_gpu": n, "iterations_per_step": iters, "point_evals_per_step": units_per_step,
                "l2": "flushed between timed steps (256 MiB fill)",
                "device_program": {"instructions": info.n_dev_instructions, "slots": info.n_slots,
                                   "constants": info.n_consts, "acc_forwards": info.n_acc_forwards},
                "sharding": "one rank per GPU, each rank its own shard, no collective",
            },
            "roofline": roofline,
            "ideal_kernel": ideal,
            "cpu_baseline": cb,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 8 * n * n_state * world, "d2h_bytes_per_step": 8 * n * n_out * world,
                    "api": "saf_iterate_host" if w.iterated else "saf_eval_host", "host_memory": "pinned"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
