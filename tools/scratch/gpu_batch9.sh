bash tools/gpu_prof.sh r02i single_expr:134217728 single_expr aizawa double_pendulum
