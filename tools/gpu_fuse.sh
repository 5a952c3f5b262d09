#!/bin/bash
# A/B of the fused-chain peephole (SAF_FUSE=0 disables it) on the shared-memory machine; optional library variants.
# Usage (under gpurun): bash tools/gpu_fuse.sh <tag> "<workloads>"
TAG=$1; WLS=$2
O=gpurun_out/$TAG; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_default.log 2>&1; echo "pytest default rc=$?" | tee -a $O/ab.txt
tail -n 2 $O/pytest_default.log | tee -a $O/ab.txt
for w in $WLS; do for f in 0 1; do
  SAF_MACHINE=s SAF_FUSE=$f timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > $O/ab.log 2>&1
  python - <<PY | tee -a $O/ab.txt
import json
l=[x for x in open("$O/ab.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$w fuse=$f", round(d["ms_per_step"],3), "ms", "%.3e"%d["value"], d["roofline"]["bound"], round(d["roofline"]["frac"],4))
else: print("$w fuse=$f FAILED", open("$O/ab.log").read()[-400:])
PY
done; done
