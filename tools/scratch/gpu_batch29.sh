#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
export TERRAIN_N=20000000
for cfg in "2 128 4 0" "2 256 2 48" "2 512 1 48" "2 512 1 24" "2 512 1 96" "2 1024 1 48" "2 256 2 16" "1 512 2 48" "1 1024 1 48" "2 512 1 200" "4 512 1 48"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 SAF_SPEC_SYNC=$4 timeout 300 python tools/spec_probe.py terrain_gradient >> gpurun_out/r02z/probe_terr2.log 2>&1
done
# pendulum: inlined leaves
for cfg in "1 64 1 100" "1 128 4 100" "2 64 1 100" "1 64 1 12"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 SAF_SPEC_INLINE_MAX=$4 timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_pend2.log 2>&1
done
cut -c1-100,180-300,360-520 gpurun_out/r02z/probe_terr2.log gpurun_out/r02z/probe_pend2.log
