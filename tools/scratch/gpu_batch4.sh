mkdir -p gpurun_out/r02d; O=gpurun_out/r02d
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -5 $O/pytest_gpu.log
bash tools/gpu_sweep.sh r02d default "terrain_gradient double_pendulum" "0"
