mkdir -p gpurun_out/r02e; O=gpurun_out/r02e
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -5 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_full.log 2> $O/bench_full.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -c 600 $O/bench_full.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1; echo "ref rc=$?" | tee -a $O/summary.txt
timeout 2400 bash tools/gpu_sanitize.sh r02e_sanitize
