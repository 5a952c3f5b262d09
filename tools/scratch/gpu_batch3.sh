mkdir -p gpurun_out/r02c; O=gpurun_out/r02c
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log
bash tools/gpu_sweep.sh r02c default "terrain_gradient" "0 2"
SAF_WIDE=0 bash tools/gpu_sweep.sh r02c_nowide default "terrain_gradient" "0"
bash tools/gpu_sweep.sh r02c default "single_expr aizawa double_pendulum clifford" "0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:saf_interp -s 4 -c 1 -o $O/prof_terrain python bench.py --workload terrain_gradient --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
echo "ncu rc=$?" | tee -a $O/summary.txt
