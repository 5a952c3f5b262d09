#!/bin/bash
# Tuning sweep: bench lines of several library builds (build/variants/*.so) x workloads.
# Usage (under gpurun): bash tools/gpu_sweep.sh <tag> "<lib names>" "<workloads>" ["P values"]
TAG=$1; LIBS=$2; WLS=$3; PS=${4:-0}
O=gpurun_out/$TAG; mkdir -p $O
for lib in $LIBS; do for w in $WLS; do for p in $PS; do
  if [ "$lib" = "default" ]; then unset SAF_B200_LIB; else export SAF_B200_LIB=$PWD/build/variants/$lib.so; fi
  if [ "$p" = "0" ]; then unset SAF_POINTS_PER_THREAD; else export SAF_POINTS_PER_THREAD=$p; fi
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > $O/sw.log 2>&1
  python - <<PY | tee -a $O/sweep.txt
import json
l=[x for x in open("$O/sw.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$lib $w P=$p", round(d["ms_per_step"],3), "%.3e"%d["value"], d["roofline"]["bound"], round(d["roofline"]["frac"],4))
else: print("$lib $w P=$p FAILED", open("$O/sw.log").read()[-300:])
PY
done; done; done
