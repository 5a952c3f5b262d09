#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
for cfg in "1 4" "1 6" "2 3" "2 5" "2 6" "2 8"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_MINB=$2 timeout 300 python tools/spec_probe.py clifford >> gpurun_out/r02z/probe_clif2.log 2>&1
done
for cfg in "4 3" "2 4" "2 8" "4 6"; do
  set -- $cfg
  SINGLE_N=134217728 SAF_SPEC_P=$1 SAF_SPEC_MINB=$2 timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_clif2.log 2>&1
done
for cfg in "1 64 1" "1 32 1" "1 128 2"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_clif2.log 2>&1
done
cut -c1-100,180-300,360-620 gpurun_out/r02z/probe_clif2.log
