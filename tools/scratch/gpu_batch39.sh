#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
for n in 12500 25000 37000 50000 75000 100000 200000; do
  PEND_N=$n SAF_SPEC_VARIANT=small timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_pendn.log 2>&1
done
PEND_N=50000 SAF_SPEC_VARIANT=small SAF_SPEC_BLOCK=32 timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_pendn.log 2>&1
PEND_N=50000 SAF_SPEC_VARIANT=small SAF_SPEC_BLOCK=64 SAF_SPEC_MINB=5 timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_pendn.log 2>&1
PEND_N=50000 SAF_SPEC_VARIANT=small SAF_SPEC_BLOCK=96 SAF_SPEC_MINB=4 timeout 300 python tools/spec_probe.py double_pendulum >> gpurun_out/r02z/probe_pendn.log 2>&1
cut -c1-100,180-300,360-660 gpurun_out/r02z/probe_pendn.log
