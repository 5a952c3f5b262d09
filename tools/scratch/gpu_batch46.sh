#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
for cfg in "2 64 1" "2 64 3" "2 128 1" "2 32 4" "1 128 3"; do
  set -- $cfg
  SAF_SPEC_VARIANT=small SAF_SPEC_P=$1 SAF_SPEC_BLOCK=$2 SAF_SPEC_MINB=$3 timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_pend5.log 2>&1
done
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_pend5.log
