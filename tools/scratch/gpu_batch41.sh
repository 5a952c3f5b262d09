#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
TERRAIN_N=20000000 timeout 600 python tools/spec_probe.py terrain_gradient double_pendulum > gpurun_out/r02z/probe_pair.log 2>&1
timeout 900 python -m pytest tests/test_gpu_specialized.py -q -x > gpurun_out/r02z/pytest_spec5.log 2>&1; echo "pytest spec rc=$?"
tail -3 gpurun_out/r02z/pytest_spec5.log
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_pair.log
