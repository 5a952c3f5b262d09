mkdir -p gpurun_out/r02b; O=gpurun_out/r02b
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log
bash tools/gpu_sweep.sh r02b default "terrain_gradient" "0 1 2 4"
SAF_WIDE=0 bash tools/gpu_sweep.sh r02b_nowide default "terrain_gradient" "2"
