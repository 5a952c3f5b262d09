mkdir -p gpurun_out/r02s; O=gpurun_out/r02s
bash tools/gpu_sweep.sh r02s default "aizawa clifford single_expr terrain_gradient double_pendulum" "0"
SAF_BALANCE=0 bash tools/gpu_sweep.sh r02s_nobal default "aizawa clifford single_expr" "0"
SAF_MACHINE=r bash tools/gpu_sweep.sh r02s_rmachine default "aizawa clifford double_pendulum" "0"
