#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
for pf in 1 0; do
  SAF_SPEC_PREFETCH=$pf SINGLE_N=134217728 timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_prefetch.log 2>&1
  SAF_SPEC_PREFETCH=$pf timeout 300 python tools/spec_probe.py single_expr clifford aizawa >> gpurun_out/r02z/probe_prefetch.log 2>&1
done
SAF_SPEC_MINB=4 SINGLE_N=134217728 timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_prefetch.log 2>&1
SAF_SPEC_MINB=6 SINGLE_N=134217728 timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_prefetch.log 2>&1
SAF_SPEC_MINB=8 SINGLE_N=134217728 timeout 300 python tools/spec_probe.py single_expr >> gpurun_out/r02z/probe_prefetch.log 2>&1
timeout 900 python -m pytest tests/test_gpu_specialized.py -q -x > gpurun_out/r02z/pytest_spec3.log 2>&1; echo "pytest spec rc=$?"
tail -3 gpurun_out/r02z/pytest_spec3.log
cut -c1-100,180-300,360-640 gpurun_out/r02z/probe_prefetch.log
