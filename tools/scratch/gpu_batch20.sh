mkdir -p gpurun_out/r02t; O=gpurun_out/r02t
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_full.log 2> $O/bench_full.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -c 400 $O/bench_full.err
