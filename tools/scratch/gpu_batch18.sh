mkdir -p gpurun_out/r02r; O=gpurun_out/r02r
bash tools/gpu_sweep.sh r02r "default nothread" "double_pendulum" "0"
SAF_POINTS_PER_THREAD=1 bash tools/gpu_sweep.sh r02r "default nothread" "terrain_gradient" "1"
