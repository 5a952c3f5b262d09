#!/bin/bash
# A/B of the two interpreter machines (SAF_MACHINE=s shared-memory register file, =r real registers) on one GPU.
# Usage (under gpurun): bash tools/gpu_ab.sh <tag> "<workloads>" [run the parity suite: 0|1]
TAG=$1; WLS=$2; TESTS=${3:-1}
O=gpurun_out/$TAG; mkdir -p $O
if [ "$TESTS" = "1" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_default.log 2>&1; echo "pytest default rc=$?" | tee -a $O/ab.txt
  SAF_MACHINE=r timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_r.log 2>&1; echo "pytest SAF_MACHINE=r rc=$?" | tee -a $O/ab.txt
  tail -3 $O/pytest_default.log $O/pytest_r.log | tee -a $O/ab.txt
fi
for w in $WLS; do for m in s r; do
  SAF_MACHINE=$m timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > $O/ab.log 2>&1
  python - <<PY | tee -a $O/ab.txt
import json
l=[x for x in open("$O/ab.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$w machine=$m", round(d["ms_per_step"],3), "ms", "%.3e"%d["value"], d["roofline"]["bound"], round(d["roofline"]["frac"],4))
else: print("$w machine=$m FAILED", open("$O/ab.log").read()[-400:])
PY
done; done
