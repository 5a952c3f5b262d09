#!/bin/bash
mkdir -p gpurun_out/r02z
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02z/pytest_pool.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02z/pytest_pool.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02z/bench_pool.log 2> gpurun_out/r02z/bench_pool.err; echo "bench rc=$?"
SAF_COPY_THREADS=10 timeout 900 python bench.py --steps 5 --warmup 3 --no-c5-sharded > gpurun_out/r02z/bench_pool10.log 2> gpurun_out/r02z/bench_pool10.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench_pool","bench_pool10"):
    d=json.loads([l for l in open(f'gpurun_out/r02z/{f}.log') if l.startswith('{')][-1])
    print(f, "clifford e2e %.3e pageable %.3e"%(d['e2e']['value'], d['e2e_pageable']['value']))
    for k,v in d['workloads'].items(): print("  ",k, "ms %.4f e2e %.3e (%.3f ms) pageable %.3e (%.3f ms)"%(v['ms_per_step'], v['e2e']['value'], v['e2e']['ms_per_step'], v['e2e_pageable']['value'], v['e2e_pageable']['ms_per_step']))
PY
