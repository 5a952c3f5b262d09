mkdir -p gpurun_out/r02k; O=gpurun_out/r02k
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -5 $O/pytest_gpu.log | cut -c1-300
for nb in 1 2 3; do
  for pts in 0 134217728; do
    SAF_STREAM_BUFS=$nb timeout 300 python bench.py --workload single_expr --points $pts --steps 10 --warmup 3 --no-cpu-baseline --no-workloads > $O/sw.log 2>&1
    python - <<PY | tee -a $O/sweep.txt
import json
l=[x for x in open("$O/sw.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("single_expr bufs=$nb points=$pts", round(d["ms_per_step"],4), "%.3e"%d["value"], d["roofline"]["bound"], round(d["roofline"]["frac"],4), "e2e %.3e"%d["e2e"]["value"])
else: print("bufs=$nb points=$pts FAILED", open("$O/sw.log").read()[-400:])
PY
  done
done
bash tools/gpu_sweep.sh r02k default "aizawa double_pendulum clifford terrain_gradient" "0"
for pts in 454656 909312; do
  timeout 300 python bench.py --workload aizawa --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > $O/sw.log 2>&1
  grep '^{' $O/sw.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('aizawa points $pts', round(d['ms_per_step'],4), '%.3e'%d['value'], round(d['roofline']['frac'],4))" | tee -a $O/sweep.txt
done
