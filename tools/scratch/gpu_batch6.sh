mkdir -p gpurun_out/r02f; O=gpurun_out/r02f
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -15 $O/pytest_gpu.log
