mkdir -p gpurun_out/r02y; O=gpurun_out/r02y
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt; free -g >> $O/gpus.txt
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.log 2> $O/bench_n$N.err; echo "bench n$N rc=$?" | tee -a $O/summary.txt
tail -c 800 $O/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref_n$N.log 2>&1; echo "ref n$N rc=$?" | tee -a $O/summary.txt
