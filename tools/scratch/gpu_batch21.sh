mkdir -p gpurun_out/r02u; O=gpurun_out/r02u
for pts in 0 134217728; do
    timeout 300 python bench.py --workload single_expr --points $pts --steps 10 --warmup 3 --no-cpu-baseline --no-workloads > $O/sw.log 2>&1
    grep '^{' $O/sw.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('single_expr points $pts', round(d['ms_per_step'],4), d.get('ms_per_step_code_warm'), round(d['roofline']['frac'],4))" | tee -a $O/sweep.txt
done
