#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_specialized.py -x -q > gpurun_out/r02z/pytest_spec.log 2>&1; echo "pytest spec rc=$?" | tee -a gpurun_out/r02z/pytest_spec.log
TERRAIN_N=20000000 timeout 600 python tools/spec_probe.py > gpurun_out/r02z/probe_default.log 2>&1
tail -5 gpurun_out/r02z/pytest_spec.log; cut -c1-100,180-300,360-560 gpurun_out/r02z/probe_default.log
