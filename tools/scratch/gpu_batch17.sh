mkdir -p gpurun_out/r02q; O=gpurun_out/r02q
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log | cut -c1-300
bash tools/gpu_sweep.sh r02q default "terrain_gradient double_pendulum aizawa clifford single_expr" "0"
