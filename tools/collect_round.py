#!/usr/bin/env python
"""Copies the judged artefacts of one tools/gpu_round.sh run from gpurun_out/<tag>/ into profiles/ under a round prefix:
bench lines, ncu launch list (+ the interpreter's share of the step), ncu summary of the Clifford capture.
Usage: python tools/collect_round.py <tag> <prefix, e.g. r01_k>"""
import csv, io, os, shutil, subprocess, sys

tag, prefix = sys.argv[1], sys.argv[2]
src, dst = os.path.join("gpurun_out", tag), "profiles"
with open(os.path.join(dst, f"{prefix}_bench_lines.md"), "w") as f:
    f.write(f"# Round 1 bench lines (one B200, tools/gpu_round.sh {tag})\n\n")
    for w in ("clifford", "single_expr", "aizawa", "double_pendulum", "terrain_gradient", "reference"):
        p = os.path.join(src, f"bench_{w}.log")
        if not os.path.exists(p):
            continue
        lines = [l for l in open(p) if l.startswith("{")]
        f.write(f"## {w}\n```json\n{lines[-1] if lines else 'no JSON line'}```\n")
    f.write("## summary\n```\n" + open(os.path.join(src, "summary.txt")).read() + "```\n")
lc = os.path.join(src, "launches.csv")
if os.path.exists(lc):
    shutil.copy(lc, os.path.join(dst, f"{prefix}_launches_clifford.csv"))
    rows = [r for r in csv.reader(l for l in open(lc) if l.startswith('"'))]
    h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    tot = {}
    for r in rows[1:]:
        tot.setdefault(r[ik], [0, 0.0]); tot[r[ik]][0] += 1; tot[r[ik]][1] += float(r[iv].replace(",", ""))
    with open(os.path.join(dst, f"{prefix}_launches_clifford.md"), "w") as f:
        f.write(f"# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` ({tag}; gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("| kernel | launches | total ms | mean ms |\n|---|---|---|---|\n")
        for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {ns / 1e6:.3f} | {ns / 1e6 / n:.3f} |\n")
        f.write("\nThe interpreter (`saf_interp_p4_ksm`) is the only kernel of the library inside a bench step (one launch per step = 1000 "
                "map iterations over 1 M particles); the other kernels are torch's L2-flush fill and state copies, the FP64-peak probe and the "
                "ideal-kernel probe, all outside the timed CUDA events.\n")
rep = os.path.join(src, "prof_clifford.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, "tools/ncu_summary.py", rep, "7812500"], capture_output=True, text=True).stdout
    stalls = subprocess.run([sys.executable, "tools/ncu_stalls.py", rep], capture_output=True, text=True).stdout
    with open(os.path.join(dst, f"{prefix}_clifford.md"), "w") as f:
        f.write(f"# Round 1 final capture: Clifford 1M x 1000 (P = 4, unified BRXU dispatch)\n\nSource: `gpurun_out/{tag}/prof_clifford.ncu-rep` "
                "(ncu --set full --clock-control none --import-source on -k regex:saf_interp -s 4 -c 1; the plain run of the same command "
                "exited 0 first).  Unit for the instruction mix: one warp-iteration (7812500 per launch).\n\n```\n" + out + "```\n\n"
                "## Sampled stall locations\n```\n" + stalls + "```\n")
print("collected", tag, "->", prefix)
