#!/bin/bash
# bench.py on all GPUs of the box under torchrun (as the driver launches it), both arms.  Usage: bash tools/gpu_multi.sh <tag>
TAG=${1:-multi}; O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt; free -g >> $O/gpus.txt
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.log 2> $O/bench_n$N.err; echo "bench n$N rc=$?" | tee -a $O/summary.txt
tail -c 600 $O/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref_n$N.log 2>&1; echo "ref n$N rc=$?" | tee -a $O/summary.txt
python - <<PY
import json
d=json.loads([l for l in open('$O/bench_n$N.log') if l.startswith('{')][-1])
print('headline', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'pageable', d['e2e_pageable']['value'], d['machine'])
print(json.dumps(d.get('c5_sharded'), indent=1)[:3000])
print('wall', d['wall_s_all_measurements'])
PY
