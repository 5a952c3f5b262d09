#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_specialized.py -q -x > gpurun_out/r02z/pytest_spec4.log 2>&1; echo "pytest spec rc=$?"
tail -3 gpurun_out/r02z/pytest_spec4.log
SINGLE_N=134217728 timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_pf2.log 2>&1
timeout 300 python tools/spec_probe.py single_expr clifford aizawa double_pendulum >> gpurun_out/r02z/probe_pf2.log 2>&1
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_pf2.log
