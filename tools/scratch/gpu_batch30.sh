#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
export TERRAIN_N=20000000
for cfg in "4 512 1 200" "4 512 1 400" "4 256 1 200" "4 256 2 200" "2 512 2 200" "2 512 1 400" "2 512 1 1000" "4 384 1 200"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 SAF_SPEC_SYNC=$4 timeout 300 python tools/spec_probe.py terrain_gradient >> gpurun_out/r02z/probe_terr3.log 2>&1
done
for cfg in "1 32 1 100" "1 128 1 100" "1 256 1 100"; do
  set -- $cfg
  SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 SAF_SPEC_INLINE_MAX=$4 timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_pend3.log 2>&1
done
# clifford / aizawa with barriers? (short loops: no) -- single_expr at 128 Mi points: HBM roofline of the specialised kernel
cut -c1-100,180-300,360-520 gpurun_out/r02z/probe_terr3.log gpurun_out/r02z/probe_pend3.log
