#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_specialized.py -q > gpurun_out/r02z/pytest_spec.log 2>&1; echo "pytest spec rc=$?" | tee -a gpurun_out/r02z/pytest_spec.log
# the whole GPU suite with every program specialised at its first launch
SAF_SPECIALIZE=1 timeout 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_register_machine.py > gpurun_out/r02z/pytest_all_spec.log 2>&1; echo "pytest all (SAF_SPECIALIZE=1) rc=$?" | tee -a gpurun_out/r02z/pytest_all_spec.log
tail -8 gpurun_out/r02z/pytest_spec.log; tail -15 gpurun_out/r02z/pytest_all_spec.log
