#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
bash tools/gpu_prof.sh r02_final3b single_expr single_expr:134217728
for cfg in "4 128 5" "4 256 2" "4 64 10" "2 128 8" "2 256 4" "4 512 1" "1 256 8"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_c1.log 2>&1
done
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_c1.log
