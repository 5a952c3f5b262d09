mkdir -p gpurun_out/r02o; O=gpurun_out/r02o
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log | cut -c1-300
bash tools/gpu_sweep.sh r02o default "double_pendulum" "0 2"
bash tools/gpu_sweep.sh r02o default "terrain_gradient" "2 1"
