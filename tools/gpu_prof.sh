#!/bin/bash
# ncu --set full capture of the library's kernel for the given workloads (one launch each), summarised ON the box:
# the .ncu-rep files are 30 MB each and gpurun_out/ travels back only below 64 MiB, so what comes back is the text
# summary (tools/ncu_summary.py: raw metrics, instruction mix per warp-iteration; tools/ncu_stalls.py: stall samples)
# and, with KEEP=1, the report of the LAST workload.
# Usage (under gpurun): bash tools/gpu_prof.sh <tag> <workload>[:points] [...]      e.g. single_expr:134217728
# MACHINE=interpreter profiles the interpreter kernel (saf_interp_*) instead of the specialised one (saf_spec, default).
TAG=$1; shift
O=gpurun_out/$TAG; mkdir -p $O
MACHINE=${MACHINE:-specialized}
KERNEL='regex:saf_spec'; [ "$MACHINE" = "interpreter" ] && KERNEL='regex:saf_r?interp'
SUF=""; [ "$MACHINE" = "interpreter" ] && SUF="_interp"
for spec in "$@"; do
  w=${spec%%:*}; pts=0; [ "$spec" != "$w" ] && pts=${spec##*:}
  name=$w$SUF; [ "$pts" != "0" ] && name=${w}${SUF}_$pts
  timeout 600 python bench.py --machine $MACHINE --workload $w --points $pts --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > $O/plain_$name.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k "$KERNEL" -s 4 -c 1 -f -o /tmp/prof_$name \
      python bench.py --machine $MACHINE --workload $w --points $pts --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > $O/ncu_$name.log 2>&1
  echo "prof $name rc=$?" | tee -a $O/summary.txt
  grep '^{' $O/plain_$name.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', '%.4f ms'%d['ms_per_step'], '%.3e'%d['value'], d['roofline']['bound'], round(d['roofline']['frac'],4), 'points', d['config']['points_per_gpu'], 'iters', d['config']['iterations_per_step'])" | tee -a $O/summary.txt
  units=$(grep '^{' $O/plain_$name.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['points_per_gpu']*d['config']['iterations_per_step']/128.0)")
  { echo "# $name ($MACHINE machine): ncu --set full --clock-control none, one launch; unit = 128 point-evals";
    python tools/ncu_summary.py /tmp/prof_$name.ncu-rep $units; python tools/ncu_stalls.py /tmp/prof_$name.ncu-rep; } > $O/ncu_summary_$name.txt 2>&1
  last=/tmp/prof_$name.ncu-rep
done
[ "$KEEP" = "1" ] && cp $last $O/
cat $O/summary.txt
