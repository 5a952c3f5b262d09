mkdir -p gpurun_out/r02p; O=gpurun_out/r02p
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log | cut -c1-300
bash tools/gpu_sweep.sh r02p default "double_pendulum aizawa clifford terrain_gradient single_expr" "0"
SAF_SUFFIX=0 bash tools/gpu_sweep.sh r02p_nosuffix default "double_pendulum aizawa" "0"
