mkdir -p gpurun_out/r02g; O=gpurun_out/r02g
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -15 $O/pytest_gpu.log | cut -c1-300
for nb in 1 2 3; do
  for pts in 0 134217728; do
    SAF_STREAM_BUFS=$nb timeout 300 python bench.py --workload single_expr --points $pts --steps 10 --warmup 3 --no-cpu-baseline > $O/sw.log 2>&1
    python - <<PY | tee -a $O/sweep.txt
import json
l=[x for x in open("$O/sw.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("single_expr bufs=$nb points=$pts", round(d["ms_per_step"],4), "%.3e"%d["value"], d["roofline"]["bound"], round(d["roofline"]["frac"],4), "e2e %.3e"%d["e2e"]["value"])
else: print("bufs=$nb points=$pts FAILED", open("$O/sw.log").read()[-400:])
PY
  done
done
