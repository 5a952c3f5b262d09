mkdir -p gpurun_out/r02l; O=gpurun_out/r02l
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -5 $O/pytest_gpu.log | cut -c1-300
timeout 1500 bash tools/gpu_checked.sh r02l_checked
