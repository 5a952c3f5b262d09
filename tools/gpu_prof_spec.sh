#!/bin/bash
# ncu --set full capture of the SPECIALISED kernel (saf_spec) of the given workloads, summarised on the box.
# Usage (under gpurun): bash tools/gpu_prof_spec.sh <tag> <workload> [...]
TAG=$1; shift
O=gpurun_out/$TAG; mkdir -p $O
for w in "$@"; do
  timeout 600 python tools/spec_probe.py $w > $O/plain_$w.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k saf_spec -s 3 -c 1 -f -o /tmp/prof_$w \
      python tools/spec_probe.py $w > $O/ncu_$w.log 2>&1
  echo "prof $w rc=$?" | tee -a $O/summary.txt
  units=$(grep '^{' $O/plain_$w.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n']*d['iters']/128.0)")
  { echo "# $w: specialised kernel, ncu --set full --clock-control none, one launch; unit = 128 point-evals";
    grep '^{' $O/plain_$w.log | tail -1 | cut -c1-600;
    python tools/ncu_summary.py /tmp/prof_$w.ncu-rep $units --hot 40; python tools/ncu_stalls.py /tmp/prof_$w.ncu-rep; } > $O/ncu_summary_$w.txt 2>&1
done
cat $O/summary.txt
