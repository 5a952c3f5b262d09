#!/bin/bash
# pendulum shape sweep (small batch: 50 k points)
mkdir -p gpurun_out/r02z
cd /root/repo
for cfg in "1 32 1" "1 32 8" "1 64 1" "1 64 4" "1 128 1" "2 32 1" "2 32 4" "2 64 1" "2 64 2" "4 32 1" "4 32 2"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_pend.log 2>&1
done
for cfg in "1 128 8" "1 256 4" "2 128 3" "2 128 5" "2 256 2"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 TERRAIN_N=20000000 timeout 300 python tools/spec_probe.py terrain_gradient >> gpurun_out/r02z/probe_terr.log 2>&1
done
for cfg in "4 128 4" "4 128 2" "4 256 2" "4 64 6" "2 128 6"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 timeout 300 python tools/spec_probe.py clifford single_expr >> gpurun_out/r02z/probe_clif.log 2>&1
done
cut -c1-420 gpurun_out/r02z/probe_pend.log gpurun_out/r02z/probe_terr.log gpurun_out/r02z/probe_clif.log
