mkdir -p gpurun_out/r02v; O=gpurun_out/r02v
bash tools/gpu_sweep.sh r02v "default sc_separate sc_noinline" "aizawa terrain_gradient clifford double_pendulum" "0"
