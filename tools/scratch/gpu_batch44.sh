#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
for pf in 1 2; do
  SAF_SPEC_PREFETCH=$pf SINGLE_N=134217728 timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_pf3.log 2>&1
  SAF_SPEC_PREFETCH=$pf SINGLE_N=134217728 SAF_SPEC_MINB=6 timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_pf3.log 2>&1
  SAF_SPEC_PREFETCH=$pf timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_pf3.log 2>&1
done
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_pf3.log
