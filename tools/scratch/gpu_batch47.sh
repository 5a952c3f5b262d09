#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
export TERRAIN_N=20000000
for cfg in "2 4" "4 3" "2 6" "4 2" "1 8"; do
  set -- $cfg
  SAF_SPEC_DEBUG=1 SAF_SPEC_P=$1 SAF_SPEC_MINB=$2 timeout 300 python tools/spec_probe.py terrain_gradient >> gpurun_out/r02z/probe_roll.log 2>&1
done
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_roll.log
