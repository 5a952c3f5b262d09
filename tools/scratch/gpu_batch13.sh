mkdir -p gpurun_out/r02m; O=gpurun_out/r02m
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.log 2> $O/bench_n2.err; echo "bench n2 rc=$?" | tee -a $O/summary.txt
tail -c 1500 $O/bench_n2.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sharded_over_devices" > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_multi.log
