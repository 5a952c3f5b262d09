#!/bin/bash
# Stand-in for compute-sanitizer memcheck (closed on this GPU pool): tools/sanitize_cases.py on the bounds-checked build
# (tools/build_variant.sh checked -DSAF_CHECKED; every shared-memory access and bulk copy of the interpreter is checked
# and traps with a message), once per forced layout.  Usage (under gpurun): bash tools/gpu_checked.sh <tag>
TAG=${1:-checked}; O=gpurun_out/$TAG; mkdir -p $O
export SAF_B200_LIB=$PWD/build/variants/checked.so
for p in 0 1 2 4; do
  if [ "$p" = "0" ]; then unset SAF_POINTS_PER_THREAD; else export SAF_POINTS_PER_THREAD=$p; fi
  for sb in 2 3; do
    SAF_STREAM_BUFS=$sb timeout 900 python tools/sanitize_cases.py > $O/checked_p${p}_s$sb.log 2>&1
    echo "checked build P=$p stream_bufs=$sb rc=$? $(grep -c 'SAF_CHECKED:' $O/checked_p${p}_s$sb.log) violations; $(tail -1 $O/checked_p${p}_s$sb.log)" | tee -a $O/summary.txt
  done
done
unset SAF_POINTS_PER_THREAD
SAF_MACHINE=r timeout 900 python tools/sanitize_cases.py > $O/checked_r.log 2>&1
echo "checked build register machine rc=$? $(grep -c 'SAF_CHECKED:' $O/checked_r.log) violations; $(tail -1 $O/checked_r.log)" | tee -a $O/summary.txt
