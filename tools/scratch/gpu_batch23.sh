mkdir -p gpurun_out/r02w; O=gpurun_out/r02w
bash tools/gpu_sweep.sh r02w "sc_halves sc_separate" "aizawa terrain_gradient clifford double_pendulum single_expr" "0"
