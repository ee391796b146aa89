#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-int8-split 2>&1 | tail -1 | tee gpurun_out/bench_n1_e2e.json | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],1), 'cpu', d['cpu_baseline']['value'])"
timeout 300 python scripts/bench_kernels.py > gpurun_out/kernels.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/kernels.json'))
for size,res in d['hbm_kernels'].items():
    print(size, {k.split(' ')[0]: round(v['GB/s']) for k,v in res.items()})
PY
