#!/bin/bash
# compute-sanitizer over the interpreter: memcheck and racecheck on tools/sanitize_cases.py, once
# per forced layout.  The kernel addresses shared memory with raw 32-bit addresses in inline PTX and its dispatch is
# rewritten at PTX level (tools/patch_brx.py): this is where an out-of-bounds slot offset or a missing barrier around
# the code window would show.  Usage (under gpurun): bash tools/gpu_sanitize.sh <tag>
TAG=${1:-sanitize}; O=gpurun_out/$TAG; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  for p in 0 1 2; do
    if [ "$p" = "0" ]; then unset SAF_POINTS_PER_THREAD; else export SAF_POINTS_PER_THREAD=$p; fi
    if [ "$tool" = "racecheck" ]; then export SAF_SANITIZE_WIDE_POINTS=610000; fi
    timeout 1500 $CS --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_cases.py > $O/${tool}_p$p.log 2>&1
    rc=$?
    echo "$tool P=$p rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_cases:' $O/${tool}_p$p.log | tr '\n' ' ')" | tee -a $O/summary.txt
  done
done
