#!/bin/bash
# Final GPU session of a round: parity suite, smoke, the full bench line (all five configs + C5 sharded), the reference
# arm, ncu launch list, and one ncu --set full capture per workload of the SHIPPED build (summaries only come back).
# Usage (under gpurun): bash tools/gpu_final.sh <tag>
TAG=${1:-final}; O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_full.log 2> $O/bench_full.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.log 2>&1; echo "reference rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a $O/summary.txt
bash tools/gpu_prof.sh $TAG clifford single_expr single_expr:134217728 aizawa double_pendulum terrain_gradient
[ "$WITH_INTERP" = "1" ] && MACHINE=interpreter bash tools/gpu_prof.sh $TAG clifford terrain_gradient
