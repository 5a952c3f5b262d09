#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
for cfg in "64 6" "64 7" "32 12" "32 14" "128 3" "96 4" "160 2" "192 2"; do
  set -- $cfg
  PEND_N=50000 SAF_SPEC_VARIANT=small SAF_SPEC_BLOCK=$1 SAF_SPEC_MINB=$2 timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_pendb.log 2>&1
done
cut -c1-100,180-300,360-660 gpurun_out/r02z/probe_pendb.log
