#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
TERRAIN_N=20000000 timeout 600 python tools/spec_probe.py > gpurun_out/r02z/probe_speculate.log 2>&1
SAF_SPEC_SPECULATE=0 timeout 600 python tools/spec_probe.py clifford double_pendulum single_expr >> gpurun_out/r02z/probe_speculate.log 2>&1
SAF_SPEC_P=2 timeout 600 python tools/spec_probe.py clifford >> gpurun_out/r02z/probe_speculate.log 2>&1
SAF_SPEC_MINB=4 timeout 600 python tools/spec_probe.py clifford >> gpurun_out/r02z/probe_speculate.log 2>&1
SAF_SPEC_MINB=2 timeout 600 python tools/spec_probe.py clifford >> gpurun_out/r02z/probe_speculate.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_specialized.py -q -x > gpurun_out/r02z/pytest_spec2.log 2>&1; echo "pytest spec rc=$?" | tee -a gpurun_out/r02z/pytest_spec2.log
tail -4 gpurun_out/r02z/pytest_spec2.log
cut -c1-100,180-300,360-620 gpurun_out/r02z/probe_speculate.log
