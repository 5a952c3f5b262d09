#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
export TERRAIN_N=20000000
for cfg in "1 400" "1 100" "0 400"; do
  set -- $cfg
  SAF_SPEC_SPECULATE_LONG=$1 SAF_SPEC_SYNC=$2 timeout 300 python tools/spec_probe.py terrain_gradient >> gpurun_out/r02z/probe_specl.log 2>&1
done
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_specl.log
