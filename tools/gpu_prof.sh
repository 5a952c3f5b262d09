#!/bin/bash
# ncu --set full capture of the interpreter kernel for the given workloads (one launch each).
# Usage (under gpurun): bash tools/gpu_prof.sh <tag> <workload> [<workload> ...]
TAG=$1; shift
O=gpurun_out/$TAG; mkdir -p $O
for w in "$@"; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_$w.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:saf_r\?interp -s 4 -c 1 -o $O/prof_$w \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_$w.log 2>&1
  echo "prof $w rc=$?" | tee -a $O/summary.txt
  grep '^{' $O/plain_$w.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'], d['roofline']['bound'], round(d['roofline']['frac'],4))" | tee -a $O/summary.txt
done
