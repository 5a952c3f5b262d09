#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
export TERRAIN_N=20000000
for cfg in "0 0" "0 1" "1 0" "1 1"; do
  set -- $cfg
  SAF_SPEC_PAIR=$1 SAF_SPEC_KTABLE=$2 timeout 600 python tools/spec_probe.py terrain_gradient >> gpurun_out/r02z/probe_pair2.log 2>&1
done
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_pair2.log
