#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU): key raw metrics + executed-instruction mix per unit of work.
Usage: python tools/ncu_summary.py <report.ncu-rep> <units: warp-iterations of the launch> [--hot N]"""
import collections, csv, io, subprocess, sys

rep, units = sys.argv[1], float(sys.argv[2])
hot_n = int(sys.argv[sys.argv.index("--hot") + 1]) if "--hot" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units_row, v = rows[0], rows[1], rows[2]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]
for k in KEYS:
    if k in h:
        print(f"{k:82s} {v[h.index(k)]}" + (f" {units_row[h.index(k)]}" if k.startswith("dram__bytes") else ""))


def _bytes(k):
    if k not in h:
        return None
    val, unit = float(v[h.index(k)].replace(",", "")), units_row[h.index(k)].lower()
    return val * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1.0)


print("DRAM_TRAFFIC_JSON " + __import__("json").dumps({"dram_bytes_read": _bytes("dram__bytes_read.sum"),
                                                        "dram_bytes_write": _bytes("dram__bytes_write.sum")}))
for i, k in enumerate(h):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            if float(v[i]) >= 0.15:
                print(f"{k:82s} {float(v[i]):.2f}")
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
iS, iE, iA, iN = h.index("Source"), h.index("Instructions Executed"), h.index("Address"), h.index("# Samples")
mix, tot, hot, base = collections.Counter(), 0, [], None
for r in rows[2:]:
    try:
        e = int(r[iE]); a = int(r[iA], 16); s = int(r[iN])
    except (ValueError, IndexError):
        continue
    base = a if base is None else base
    if not e:
        continue
    t = r[iS].split()
    o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    mix[o] += e; tot += e
    hot.append((a - base, e / units, s, r[iS].strip()))
print(f"executed warp-instructions per unit: {tot / units:.1f}")
fp64 = sum(mix[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"  FP64 {fp64 / units:.1f} ({100 * fp64 / tot:.1f} %)")
for o, e in mix.most_common(18):
    print(f"  {o:10s} {e / units:8.1f}  {100 * e / tot:5.1f} %")
if hot_n:
    for a, c, s, t in hot:
        if c >= 0.5:
            print(f"{a:6x} {c:7.2f} {s:6d} {t}")
