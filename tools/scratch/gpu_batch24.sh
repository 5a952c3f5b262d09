mkdir -p gpurun_out/r02x; O=gpurun_out/r02x
bash tools/gpu_sweep.sh r02x "default pad2 pad4 pad6" "aizawa clifford double_pendulum terrain_gradient single_expr" "0"
