#!/usr/bin/env python
"""bench.py -- f64 point-evals/s of the SymbAnaFis batch evaluator on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input.  Headline workload = BASELINE.json
configs[1]: the Clifford attractor map, 1M particles x 1000 iterations per GPU, i.e. one step = 1e9 point-evals per
GPU (one point-eval = the whole 2-output program on one particle).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

The ONE JSON line rank 0 prints carries, besides the headline workload at the top level:
  "workloads"  : the other four BASELINE.json configs (single_expr, aizawa, double_pendulum, terrain_gradient), each
                 with ms_per_step, value, roofline{bound, frac, traffic}, e2e, e2e_pageable, cpu_baseline  (N = 1)
  "c5_sharded" : BASELINE.json configs[4] as north_star states it -- 100 M points of ONE seeded global array cut into N
                 contiguous shards, one rank per GPU through saf_eval_host, and once from rank 0 alone through
                 saf_eval_host_multi (one host thread + stream set per device, host gather included)

N > 1 is launched by torchrun (one rank per GPU); points are sharded per rank, there is no collective on the data path
(SURVEY.md section 8e), every timing is the max over ranks.

`value`        : device-timed (CUDA events on the launching stream), columns resident in HBM, L2 flushed between steps.
`e2e`          : the same step through the C-ABI host call (saf_iterate_host / saf_eval_host) with pinned HOST buffers:
                 H2D of the inputs and D2H of the results inside the timed region.
`e2e_pageable` : the same with ordinary (pageable) NumPy arrays -- what a caller of the reference holds.
`machine`      : which of the library's two machines ran the program (--machine): "specialized" (default) = the
                 program's own straight-line sm_100a kernel, compiled once by saf_program_specialize before the timed
                 region -- like CompiledEvaluator::compile, outside every timed region; its cost is reported in
                 machine.compile_ms -- or "interpreter" = the bytecode interpreter kernel, which needs no compilation.
                 Both return the same bits; "interpreter" beside each workload is the interpreter's device time.
`--impl reference`: the reference's CPU algorithm (oracle port of the f64x4 + 256-point-chunk path, all host threads)
                 on a bounded sample of the same workload; same `config` object, byte for byte.
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
HEADLINE = "clifford"
OTHER_WORKLOADS = ("single_expr", "aizawa", "double_pendulum", "terrain_gradient")
C5_TOTAL_POINTS = 100_000_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--points", type=int, default=0, help="points per GPU (0 = BASELINE.json size)")
    ap.add_argument("--iters", type=int, default=0, help="map iterations per step (0 = BASELINE.json value)")
    ap.add_argument("--machine", default="specialized", choices=["specialized", "interpreter"],
                    help="specialized: saf_program_specialize before the timed region (default); interpreter: never")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="headline workload only (profiling runs)")
    ap.add_argument("--no-c5-sharded", action="store_true")
    ap.add_argument("--c5-points", type=int, default=C5_TOTAL_POINTS, help="total points of the sharded C5 run")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe)."""

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


def measured_traffic(workload_name, machine="interpreter"):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernel launch from the committed ncu captures of the shipped
    build (profiles/r02_dram_traffic_specialized.json for the specialised kernels, profiles/r02_dram_traffic.json for the
    interpreter), or None when that workload has not been captured."""
    names = ("r02_dram_traffic_specialized.json",) if machine == "specialized" else ("r02_dram_traffic.json", "r01_dram_traffic.json")
    for name in names:
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f).get(workload_name)
            if t is not None:
                return t["dram_bytes_read"] + t["dram_bytes_write"], f"profiles/{name}"
        except (OSError, ValueError, KeyError):
            continue
    return None, None


def workload_sizes(w, args, headline: bool):
    n = (args.points if headline and args.points else w.n_points)
    iters = (args.iters if headline and args.iters else (w.iterations if w.iterated else 1))
    return n, iters


def config_for(w, n, iters, world):
    """The workload's identity -- the SAME object in the GPU arm and in the reference arm."""
    return {
        "workload": f"{w.name}: {w.description}",
        "points_per_gpu": n, "iterations_per_step": iters, "point_evals_per_step": n * iters * world,
        "l2": "GPU arm: flushed between timed steps (256 MiB fill)",
        "sharding": "one rank per GPU, each rank its own shard, no collective",
    }


# ---- the CPU legs (oracle: test infrastructure; timed here as the reported baseline, never the product) -------------
def cpu_leg(w, n, iters, budget_s, steps=1, warmup=0):
    """The reference's eval_f64 path restated (f64x4 body, 256-point chunks, one workspace per thread:
    oracle/saf_oracle_simd.c) on all host threads, on a bounded sample of the workload."""
    import oracle
    from oracle import OracleProgram
    op = OracleProgram.from_program(w.program)
    cores = oracle.online_cores()
    n_s = min(n, 1_000_000)
    cols = w.make_inputs(n_s, 1)
    t0 = time.perf_counter()  # calibration
    if w.iterated:
        op.iterate(cols, 2, cores, simd=True)
    else:
        op.run_chunked_simd(cols, n_s, cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    per_unit = dt / (n_s * (2 if w.iterated else 1))
    it_s = int(max(2, min(iters, budget_s / max(per_unit * n_s, 1e-9)))) if w.iterated else 1
    reps = 1 if w.iterated else int(max(1, min(200, budget_s / max(per_unit * n_s, 1e-9))))
    state = [c.copy() for c in cols]

    def step():
        if w.iterated:
            for s, c in zip(state, cols):
                s[:] = c
            op.iterate(state, it_s, cores, simd=True, in_place=True)
        else:
            for _ in range(reps):
                op.run_chunked_simd(cols, n_s, cores)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    units = n_s * (it_s if w.iterated else reps) * steps
    sample = (f"{n_s} points x {it_s} iterations per step" if w.iterated else f"{reps} passes over {n_s} points per step")
    kind = "port-f64x4" if oracle.has_simd() else "port"
    return {"value": units / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + f" of {w.name}",
            "seconds": round(dt, 3)}, dt / steps


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from symbanafis_b200 import workloads
    w = workloads.get(args.workload)
    n, iters = workload_sizes(w, args, True)
    cb, step_s = cpu_leg(w, n, iters, budget_s=2.0, steps=args.steps, warmup=args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(w, n, iters, max(1, args.gpus)),
        "sample": cb["sample"],
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference's eval_f64 path (oracle/: f64x4 body + 256-point chunks, all host "
                "threads); the Rust reference cannot be built in this image (no cargo/rustc)",
    }))


# ---- the GPU arm -----------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from symbanafis_b200 import _lib
        self.torch, self.dist, self._lib, self.args = torch, dist, _lib, args
        self.rank, self.world, self.local = dist_env()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the evaluator has no CPU path")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            # a barrier that keeps the waiting ranks' GPUs idle (an NCCL barrier spins in a kernel): used while rank 0
            # alone drives all devices through saf_eval_host_multi
            self.cpu_group = dist.new_group(backend="gloo")
        self.machine_note = None
        if args.machine == "specialized" and not _lib.specialize_available():
            # no usable libnvrtc on this box: the interpreter runs everything (it always can); say so in the line
            self.machine_note = "specialised kernels unavailable (" + _lib.lib().saf_last_error().decode() + "): interpreter"
            args.machine = "interpreter"
        self.flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=self.dev)  # 256 MiB > 126 MB L2
        self.stream = torch.cuda.current_stream(self.dev)
        self.fp64_peak = _lib.measure_fp64_peak()  # DFMA lane-instr/s, measured live on this GPU
        self.hbm_gbs, self.peak_src = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, vals):
        from symbanafis_b200.sharding import max_over_ranks
        return max_over_ranks(vals, device=self.dev)

    def handle(self, w, machine=None):
        p = w.program
        h = self._lib.DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size,
                                                 p.result_regs)
        if (machine or self.args.machine) == "specialized":
            h.specialize()  # NVRTC, once per program, outside every timed region (reported: machine.compile_ms)
        return h

    def machine_info(self, h):
        if self.args.machine != "specialized":
            out = {"name": "interpreter", "kernel": "saf_interp_* (bytecode interpreter, shared-memory register file)"}
            if self.machine_note:
                out["note"] = self.machine_note
            return out
        si = h.specialized_info()
        small = si.small_registers > 0 and si.registers <= 0
        return {"name": "specialized", "kernel": "saf_spec (the program as straight-line sm_100a code, NVRTC)",
                "shape": "small-batch" if small else "large-batch",
                "points_per_thread": si.small_points_per_thread if small else si.points_per_thread,
                "block_threads": si.small_block_threads if small else si.block_threads,
                "registers": si.small_registers if small else si.registers,
                "compile_ms": si.compile_ms + (si.small_compile_ms if si.small_compiled else 0.0),
                "compile_note": "one-off, before the timed region (saf_program_specialize); the interpreter needs none"}

    # one workload, device-resident columns: K timed steps, CUDA events on the launching stream
    def device_timed(self, h, w, cols_np, n, iters, steps, warmup, extras=True):
        torch = self.torch
        cols0 = [torch.from_numpy(c).to(self.dev) for c in cols_np]
        n_out = len(w.program.result_regs)
        state = [c.clone() for c in cols0] if w.iterated else None
        outs = None if w.iterated else [torch.empty(n, dtype=torch.float64, device=self.dev) for _ in range(n_out)]
        s = self.stream.cuda_stream

        def launch():
            if w.iterated:
                h.iterate_device([t.data_ptr() for t in state], n, iters, stream=s)
            else:
                h.eval_device([c.data_ptr() for c in cols0], [n] * len(cols0), [o.data_ptr() for o in outs], n, stream=s)

        def reset():
            if w.iterated:
                for t, c in zip(state, cols0):
                    t.copy_(c)

        for _ in range(max(3, warmup)):
            reset()
            launch()
        self.barrier()
        launches0 = self._lib.kernel_launch_count()
        ms = []
        t_wall0 = time.perf_counter()
        for _ in range(steps):
            self.flush.fill_(1.0)  # evict L2 between timed steps (not timed)
            reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            launch()
            e1.record(self.stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        self.barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = self._lib.kernel_launch_count() - launches0
        # the same measurement with the CODE warm: the L2 flush also evicts the kernel's instructions and the program
        # image, which a launch of a few microseconds then fetches from DRAM (ncu: no_instruction is the top stall of
        # the 1 M-point single pass).  A 512-point launch of the same program on scratch columns after the flush
        # brings code and image back without touching the timed data; reported beside ms_per_step, never instead.
        warm_ms = None
        if not w.iterated and steps > 0 and extras:
            # (the specialised machine picks its kernel by batch size: the scratch launch must be large enough to run
            # the SAME kernel as the timed launch -- 256 Ki points, 4 MB of L2)
            m_ = 512 if self.args.machine == "interpreter" else min(n, 256 * 1024)
            scr_in = [torch.zeros(m_, dtype=torch.float64, device=self.dev) for _ in cols0]
            scr_out = [torch.empty(m_, dtype=torch.float64, device=self.dev) for _ in range(n_out)]
            wm = []
            for _ in range(steps):
                self.flush.fill_(1.0)
                h.eval_device([c.data_ptr() for c in scr_in], [m_] * len(scr_in), [o.data_ptr() for o in scr_out], m_, stream=s)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.stream)
                launch()
                e1.record(self.stream)
                e1.synchronize()
                wm.append(e0.elapsed_time(e1))
            warm_ms = sum(wm) / len(wm)
        # sustained: back-to-back steps for >= 2 s (no flush, no host sync in between): clocks under a long load
        sustained = None
        if w.name == HEADLINE and not self.args.no_workloads and extras:
            per = max(sum(ms) / len(ms), 1e-3)
            k = int(min(2000, max(10, 2200.0 / per)))
            reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(k):
                launch()
            e1.record(self.stream)
            e1.synchronize()
            t = e0.elapsed_time(e1)
            sustained = {"steps": k, "seconds": t * 1e-3, "ms_per_step": t / k,
                         "note": "back-to-back launches, no L2 flush or host sync in between"}
        return sum(ms) / len(ms), t_wall, launches, sustained, warm_ms

    # the host-buffer C-ABI call: H2D of the inputs and D2H of the results inside the timed region
    def e2e_timed(self, h, w, cols_np, n, iters, steps, pinned: bool):
        torch = self.torch
        n_out = len(w.program.result_regs)
        if pinned:
            t_in = [torch.from_numpy(c).pin_memory() for c in cols_np]
            ins = [t.numpy() for t in t_in]
            t_state = [torch.empty_like(t).pin_memory() for t in t_in] if w.iterated else None
            t_out = None if w.iterated else [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(n_out)]
            state = [t.numpy() for t in t_state] if w.iterated else None
            outs = None if w.iterated else [t.numpy() for t in t_out]
        else:
            ins = [np.array(c, copy=True) for c in cols_np]
            state = [np.empty_like(c) for c in ins] if w.iterated else None
            outs = None if w.iterated else [np.empty(n) for _ in range(n_out)]

        def step(timed):
            if w.iterated:
                for s_, c in zip(state, ins):
                    s_[:] = c
            t0 = time.perf_counter()
            if w.iterated:
                h.iterate_host(state, iters)
            else:
                h.eval_host(ins, outs, n)
            return time.perf_counter() - t0

        for _ in range(2):
            step(False)
        self.barrier()
        ts = [step(True) for _ in range(steps)]
        self.barrier()
        return 1e3 * sum(ts) / len(ts)

    def roofline(self, info, w, n, iters, step_ms, full_size, machine=None):
        bytes_launch = info.bytes_per_point * n               # columns read + written once per launch
        slots_launch = info.fp64_slots * n * iters            # FP64 issue slots per launch
        t_hbm = bytes_launch / (self.hbm_gbs * 1e9)
        t_fp64 = slots_launch / self.fp64_peak if self.fp64_peak > 0 else 0.0
        dur = step_ms * 1e-3
        traffic, traffic_src = measured_traffic(w.name, machine or self.args.machine) if full_size else (None, None)
        if t_fp64 >= t_hbm:
            achieved, peak = 2.0 * slots_launch / dur / 1e12, 2.0 * self.fp64_peak / 1e12
            return {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak if peak else None, "traffic": traffic,
                    "traffic_unit": f"DRAM bytes per launch (ncu --set full capture, {traffic_src})",
                    "peak_source": "DFMA-chain probe measured in this run (saf_measure_fp64_peak); "
                                   "FP64 is not in MEASURED_PEAKS.json",
                    "algorithmic": f"{info.fp64_slots:g} FP64 issue slots/point-eval (FMA = 2 FLOP) x "
                                   f"{n * iters} point-evals/launch",
                    "hbm_leg": {"achieved_gbs": bytes_launch / dur / 1e9, "peak_gbs": self.hbm_gbs,
                                "peak_source": self.peak_src}}
        achieved = bytes_launch / dur / 1e9
        return {"bound": "hbm", "achieved": achieved, "peak": self.hbm_gbs, "unit": "GB/s",
                "frac": achieved / self.hbm_gbs, "traffic": traffic,
                "traffic_unit": f"DRAM bytes per launch (ncu --set full capture, {traffic_src})",
                "peak_source": f"MEASURED_PEAKS.json ({self.peak_src})",
                "algorithmic": f"{info.bytes_per_point:g} B/point x {n} points/launch",
                "fp64_leg": {"achieved_tflops": 2.0 * slots_launch / dur / 1e12,
                             "peak_tflops": 2.0 * self.fp64_peak / 1e12}}

    def measure(self, w, n, iters, steps, warmup, seed, cpu_budget_s, headline):
        """Everything for one workload; returns the dict rank 0 prints (None on other ranks)."""
        h = self.handle(w)
        info = h.info()
        n_state, n_out = w.program.param_count, len(w.program.result_regs)
        cols_np = w.make_inputs(n, seed)
        step_ms, t_wall, launches, sustained, warm_ms = self.device_timed(h, w, cols_np, n, iters, steps, warmup)
        if step_ms < 0.1 and steps < 50:
            # a launch of a few microseconds: one hiccup of the box (seen once: 0.18 ms among 0.010 ms steps) would
            # decide the mean of a handful of steps -- time 50 of them instead (still the mean, still every step flushed)
            steps = 50
            step_ms, t_wall, launches, sustained, warm_ms = self.device_timed(h, w, cols_np, n, iters, steps, warmup)
        e2e_steps = max(2, min(steps, 5))
        e2e_ms = self.e2e_timed(h, w, cols_np, n, iters, e2e_steps, pinned=True)
        e2e_pg_ms = self.e2e_timed(h, w, cols_np, n, iters, e2e_steps, pinned=False)
        interp_ms = None
        if self.args.machine == "specialized":  # the interpreter on the same columns, device-timed, for comparison
            hi = self.handle(w, "interpreter")
            interp_ms = self.device_timed(hi, w, cols_np, n, iters, max(2, min(steps, 3)), 3, extras=False)[0]
            interp_ms = self.max_over_ranks([interp_ms])[0]
        step_ms, e2e_ms, e2e_pg_ms, t_wall = self.max_over_ranks([step_ms, e2e_ms, e2e_pg_ms, t_wall])
        if self.rank != 0:
            return None
        units = n * iters * self.world
        api = "saf_iterate_host" if w.iterated else "saf_eval_host"
        io = {"h2d_bytes_per_step": 8 * n * n_state * self.world, "d2h_bytes_per_step": 8 * n * n_out * self.world, "api": api}
        out = {
            "ms_per_step": step_ms, "value": units / (step_ms * 1e-3), "unit": UNIT, "steps": steps,
            "config": config_for(w, n, iters, self.world),
            "device_program": {"instructions": info.n_dev_instructions, "slots": info.n_slots,
                               "constants": info.n_consts, "acc_forwards": info.n_acc_forwards},
            "roofline": self.roofline(info, w, n, iters, step_ms, n == w.n_points),
            "machine": self.machine_info(h),
            "e2e": dict(value=units / (e2e_ms * 1e-3), unit=UNIT, ms_per_step=e2e_ms, host_memory="pinned", **io),
            "e2e_pageable": dict(value=units / (e2e_pg_ms * 1e-3), unit=UNIT, ms_per_step=e2e_pg_ms,
                                 host_memory="pageable (NumPy arrays, staged through the library's pinned ring)", **io),
            "gpu_launches": int(launches), "wall_s_timed_region": t_wall,
        }
        if interp_ms is not None:
            out["interpreter"] = {"ms_per_step": interp_ms, "value": units / (interp_ms * 1e-3), "unit": UNIT,
                                  "roofline_frac": self.roofline(info, w, n, iters, interp_ms, False)["frac"],
                                  "note": "the same program and columns on the bytecode interpreter kernel (same bits)"}
        if sustained:
            out["sustained"] = sustained
        if warm_ms is not None:
            out["ms_per_step_code_warm"] = warm_ms
            out["code_warm_note"] = ("data cold (L2 flushed), kernel code and program image warm (a 512-point launch of the "
                                     "same program on scratch columns between the flush and the timed launch)")
        if headline and w.name == "clifford":
            # the same map hard-coded around the same sin/cos kernels, no interpreter: what interpretation costs
            t_ideal = self._lib.measure_clifford_probe(n, iters)
            out["ideal_kernel"] = {"kernel": "saf_clifford_probe_kernel (hard-coded map, state in registers, same launch shape)",
                                   "ms_per_step": 1e3 * t_ideal, "value": n * iters / t_ideal, "unit": UNIT}
        out["cpu_baseline"] = None if self.args.no_cpu_baseline else cpu_leg(w, n, iters, cpu_budget_s)[0]
        return out

    # ---- BASELINE.json configs[4] as north_star states it: 100 M points sharded contiguously over the GPUs ----
    def c5_sharded(self, total):
        from symbanafis_b200 import workloads
        from symbanafis_b200.sharding import shard_bounds
        torch = self.torch
        w = workloads.get("terrain_gradient")
        h = self.handle(w)
        info = h.info()
        n_out = len(w.program.result_regs)
        glob = w.make_inputs(total, 42)                      # ONE seeded global array, the same for every N
        b, e = shard_bounds(total, self.world, self.rank)    # the split of saf_eval_host_multi and of batch.rs:55-74
        n = e - b
        shard = [np.ascontiguousarray(c[b:e]) for c in glob]
        # device-timed shard
        cols_d = [torch.from_numpy(c).to(self.dev) for c in shard]
        outs_d = [torch.empty(n, dtype=torch.float64, device=self.dev) for _ in range(n_out)]
        s = self.stream.cuda_stream
        ms = []
        for i in range(5):
            self.flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            h.eval_device([c.data_ptr() for c in cols_d], [n] * len(cols_d), [o.data_ptr() for o in outs_d], n, stream=s)
            e1.record(self.stream)
            e1.synchronize()
            if i >= 2:
                ms.append(e0.elapsed_time(e1))
        kernel_ms = sum(ms) / len(ms)
        # every rank its shard through saf_eval_host (pinned and pageable host buffers), all ranks at once
        pin_in = [torch.from_numpy(c).pin_memory() for c in shard]
        pin_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(n_out)]
        pg_out = [np.empty(n) for _ in range(n_out)]

        def timed(fn, reps=3):
            fn()
            ts = []
            for _ in range(reps):
                self.barrier()
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return 1e3 * sum(ts) / len(ts)

        pinned_ms = timed(lambda: h.eval_host([t.numpy() for t in pin_in], [t.numpy() for t in pin_out], n))
        pageable_ms = timed(lambda: h.eval_host(shard, pg_out, n))
        same = all(np.array_equal(a.numpy().view(np.uint64), b_.view(np.uint64)) for a, b_ in zip(pin_out, pg_out))
        # what the links alone deliver with every rank copying at once: names the limiter of the end-to-end numbers
        d_tmp = torch.empty(n, dtype=torch.float64, device=self.dev)

        def copies():
            for t in pin_in:
                d_tmp.copy_(t, non_blocking=True)
            for t in pin_out:
                t.copy_(d_tmp, non_blocking=True)
            torch.cuda.synchronize(self.dev)

        copy_ms = timed(copies)
        kernel_ms, pinned_ms, pageable_ms, copy_ms = self.max_over_ranks([kernel_ms, pinned_ms, pageable_ms, copy_ms])
        # once more from ONE process: rank 0 drives all N devices through saf_eval_host_multi (host gather included)
        multi = None
        self.barrier()
        del cols_d, outs_d, d_tmp
        torch.cuda.empty_cache()
        if self.rank == 0:
            outs = [np.empty(total) for _ in range(n_out)]
            devs = list(range(self.world))
            h.eval_host(glob, outs, total, devices=devs)     # warm-up: contexts, modules, staging rings
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                h.eval_host(glob, outs, total, devices=devs)
                ts.append(time.perf_counter() - t0)
            t_multi = sum(ts) / len(ts)
            ok = all(np.array_equal(o[b:e].view(np.uint64), p_.view(np.uint64)) for o, p_ in zip(outs, pg_out))
            # the same call on pinned host arrays (a caller that registered its buffers): no staging copies on the host
            g_pin = [torch.from_numpy(c).pin_memory() for c in glob]
            o_pin = [torch.empty(total, dtype=torch.float64).pin_memory() for _ in range(n_out)]
            gp, op_ = [t.numpy() for t in g_pin], [t.numpy() for t in o_pin]
            h.eval_host(gp, op_, total, devices=devs)
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                h.eval_host(gp, op_, total, devices=devs)
                ts.append(time.perf_counter() - t0)
            t_multi_pin = sum(ts) / len(ts)
            ok_pin = all(np.array_equal(a.view(np.uint64), b_.view(np.uint64)) for a, b_ in zip(op_, outs))
            multi = {"api": "saf_eval_host_multi (one process, one host thread + stream set per device, host gather)",
                     "devices": self.world,
                     "pageable": {"ms": 1e3 * t_multi, "value": total / t_multi, "unit": UNIT,
                                  "note": "3.2 GB staged through the pinned ring by host copy threads: host-memory bound"},
                     "pinned": {"ms": 1e3 * t_multi_pin, "value": total / t_multi_pin, "unit": UNIT},
                     "rank0_shard_bits_equal_per_rank_run": bool(ok), "pinned_bits_equal_pageable": bool(ok_pin)}
        if self.cpu_group is not None:
            self.dist.barrier(group=self.cpu_group)  # the other ranks wait here on the CPU, GPUs idle
        self.barrier()
        if self.rank != 0:
            return None
        bytes_total = 8 * total * (len(glob) + n_out)
        t_k, t_c = kernel_ms * 1e-3, copy_ms * 1e-3
        limiter = (f"kernel (FP64-bound, {self.args.machine} machine)" if t_k >= t_c else
                   "host<->device copies (PCIe / host DRAM shared by all ranks), not a collective: there is none")
        roof = self.roofline(info, w, n, 1, kernel_ms, False)
        return {
            "config": f"terrain_gradient: {total} points of one seeded global array, {self.world} contiguous shard(s), "
                      "program replicated, no collective (host gather only)",
            "total_points": total, "points_per_gpu": n, "n_gpus": self.world, "scaling": "strong",
            "kernel": {"ms": kernel_ms, "value": total / t_k, "unit": UNIT, "roofline_frac": roof["frac"], "bound": roof["bound"]},
            "e2e": {"api": "saf_eval_host per rank", "host_memory": "pinned", "ms": pinned_ms,
                    "value": total / (pinned_ms * 1e-3), "unit": UNIT},
            "e2e_pageable": {"api": "saf_eval_host per rank", "host_memory": "pageable", "ms": pageable_ms,
                             "value": total / (pageable_ms * 1e-3), "unit": UNIT,
                             "bits_equal_pinned_run": bool(same)},
            "host_multi": multi,
            "machine": self.machine_info(h),
            "copies_only": {"ms": copy_ms, "gbs_aggregate": bytes_total / t_c / 1e9,
                            "note": "the same H2D + D2H bytes with pinned buffers and no kernel, all ranks at once"},
            "h2d_bytes": 8 * total * len(glob), "d2h_bytes": 8 * total * n_out,
            "limiter": limiter,
        }


def run_ours(args):
    from symbanafis_b200 import workloads
    B = Bench(args)
    w = workloads.get(args.workload)
    n, iters = workload_sizes(w, args, True)
    sampler = ClockSampler(B.local)
    if B.rank == 0:
        sampler.start()
    t_all0 = time.perf_counter()
    main = B.measure(w, n, iters, args.steps, args.warmup, 1 + B.rank, 12.0, True)
    others = {}
    full_run = args.workload == HEADLINE and not args.no_workloads and not args.points and not args.iters
    if full_run and B.world == 1:
        for name in OTHER_WORKLOADS:
            ww = workloads.get(name)
            nn, ii = workload_sizes(ww, args, False)
            others[name] = B.measure(ww, nn, ii, max(3, min(args.steps, 10)), 3, 1, 4.0, False)
    big = None
    if full_run and B.world == 1:
        # the single pass of C1 where it IS bandwidth-bound: 128 Mi points (1 GiB read + 1 GiB written), device-timed
        ww = workloads.get("single_expr")
        nn = 128 * 1024 * 1024
        hh = B.handle(ww)
        ms_, _, _, _, warm_ = B.device_timed(hh, ww, ww.make_inputs(nn, 1), nn, 1, 5, 3)
        big = {"config": config_for(ww, nn, 1, 1), "ms_per_step": ms_, "value": nn / (ms_ * 1e-3), "unit": UNIT,
               "roofline": B.roofline(hh.info(), ww, nn, 1, ms_, False), "machine": B.machine_info(hh),
               "note": "BASELINE.json configs[0] at 128 Mi points instead of 1 M: the streaming path (bulk-copied column tiles)"}
    c5 = None
    if full_run and not args.no_c5_sharded:
        if B.world == 1 and "terrain_gradient" in others and args.c5_points == C5_TOTAL_POINTS:
            t = others["terrain_gradient"]  # one GPU: the shard IS the whole array (same seed-42 generator, 100 M points)
            c5 = {"config": f"terrain_gradient: {C5_TOTAL_POINTS} points, 1 shard (= workloads.terrain_gradient)",
                  "total_points": C5_TOTAL_POINTS, "points_per_gpu": C5_TOTAL_POINTS, "n_gpus": 1, "scaling": "strong",
                  "kernel": {"ms": t["ms_per_step"], "value": t["value"], "unit": UNIT,
                             "roofline_frac": t["roofline"]["frac"], "bound": t["roofline"]["bound"]},
                  "e2e": {"api": "saf_eval_host", "host_memory": "pinned", "ms": t["e2e"]["ms_per_step"],
                          "value": t["e2e"]["value"], "unit": UNIT},
                  "e2e_pageable": {"api": "saf_eval_host", "host_memory": "pageable", "ms": t["e2e_pageable"]["ms_per_step"],
                                   "value": t["e2e_pageable"]["value"], "unit": UNIT},
                  "host_multi": None}
        else:
            c5 = B.c5_sharded(args.c5_points)
    t_all = time.perf_counter() - t_all0
    clocks = sampler.stop() if B.rank == 0 else None
    if B.rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": B.world, "steps": main["steps"],
            "warmup": max(3, args.warmup), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": main["config"],
            "device_program": main["device_program"],
            "roofline": main["roofline"],
            "machine": main["machine"],
            "interpreter": main.get("interpreter"),
            "ideal_kernel": main.get("ideal_kernel"),
            "sustained": main.get("sustained"),
            "cpu_baseline": main["cpu_baseline"],
            "e2e": main["e2e"],
            "e2e_pageable": main["e2e_pageable"],
            "gpu_launches": main["gpu_launches"],
            "clocks": clocks,
            "wall_s_timed_region": main["wall_s_timed_region"],
            "wall_s_all_measurements": t_all,
        }
        if others:
            line["workloads"] = others
        if big is not None:
            line["single_expr_128Mi"] = big
        if c5 is not None:
            line["c5_sharded"] = c5
        print(json.dumps(line))
    if B.world > 1:
        B.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
