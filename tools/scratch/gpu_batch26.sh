#!/bin/bash
# first run of the specialised kernels
mkdir -p gpurun_out/r02z
cd /root/repo
timeout 600 python tools/spec_probe.py > gpurun_out/r02z/probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/r02z/probe.log
for P in 1 2 4; do
  SAF_SPEC_P=$P timeout 300 python tools/spec_probe.py aizawa double_pendulum clifford >> gpurun_out/r02z/probe_p.log 2>&1
done
SAF_SPEC_P=1 TERRAIN_N=20000000 timeout 300 python tools/spec_probe.py terrain_gradient >> gpurun_out/r02z/probe_p.log 2>&1
SAF_SPEC_P=4 timeout 300 python tools/spec_probe.py terrain_gradient >> gpurun_out/r02z/probe_p.log 2>&1
tail -n 30 gpurun_out/r02z/probe.log gpurun_out/r02z/probe_p.log
