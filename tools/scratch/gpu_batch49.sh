#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
export TERRAIN_N=20000000 SAF_SPEC_REROLL=1
for cfg in "4 3" "2 5" "4 2"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_MINB=$2 timeout 300 python tools/spec_probe.py terrain_gradient >> gpurun_out/r02z/probe_roll2.log 2>&1
done
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_roll2.log
