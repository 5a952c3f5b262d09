#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
timeout 1500 python bench.py > gpurun_out/r02z/bench_spec.log 2> gpurun_out/r02z/bench_spec.err; echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
tail -3 gpurun_out/r02z/bench_spec.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02z/bench_spec.log') if l.startswith('{')][-1])
def row(name,x):
    print(name, "ms %.4f"%x["ms_per_step"], "value %.3e"%x["value"], x["roofline"]["bound"], "frac %.3f"%x["roofline"]["frac"], "e2e %.3e"%x["e2e"]["value"], "pageable %.3e"%x["e2e_pageable"]["value"], "interp ms %.4f"%x["interpreter"]["ms_per_step"], x["machine"].get("shape"), x["machine"].get("points_per_thread"), x["machine"].get("registers"), "compile %.0f ms"%x["machine"].get("compile_ms",0))
row("clifford", d)
for k,v in d["workloads"].items(): row(k,v)
print("128Mi", d["single_expr_128Mi"]["ms_per_step"], d["single_expr_128Mi"]["roofline"]["frac"])
print("sustained", d["sustained"]); print("wall", d["wall_s_all_measurements"]); print("clocks", d["clocks"]); print("ideal", d["ideal_kernel"])
PY
