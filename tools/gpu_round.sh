#!/bin/bash
# One GPU session: parity suite, smoke, bench lines for every named workload, ncu launch list + full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [profile:0|1]
TAG=${1:-run}; PROF=${2:-1}
mkdir -p gpurun_out/$TAG
O=gpurun_out/$TAG
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
for w in clifford single_expr aizawa double_pendulum terrain_gradient; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > $O/bench_$w.log 2>&1; echo "bench $w rc=$?" | tee -a $O/summary.txt
  grep '^{' $O/bench_$w.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'], d['roofline']['bound'], round(d['roofline']['frac'],4), 'e2e %.3e'%d['e2e']['value'], 'cpu', d['cpu_baseline'] and '%.3e'%d['cpu_baseline']['value'])" | tee -a $O/summary.txt
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.log 2>&1; echo "bench reference rc=$?" | tee -a $O/summary.txt
if [ "$PROF" = "1" ]; then
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
  echo "ncu launches rc=$?" | tee -a $O/summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:saf_interp -s 4 -c 1 -o $O/prof_clifford python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
  echo "ncu full rc=$?" | tee -a $O/summary.txt
fi
cat $O/summary.txt
