mkdir -p gpurun_out/r02a; O=gpurun_out/r02a
nvidia-smi > $O/smi.txt 2>&1
build/ubench > $O/ubench.txt 2>&1; echo "ubench rc=$?" >> $O/summary.txt
bash tools/gpu_sweep.sh r02a default "terrain_gradient" "0 1 2 4"
SAF_WIDE=0 bash tools/gpu_sweep.sh r02a_nowide default "terrain_gradient" "1 2"
bash tools/gpu_sweep.sh r02a default "single_expr aizawa double_pendulum clifford" "0"
