#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --size 65536 --grid 4 --dtype bf16 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_n1.json | cut -c1-330
timeout 300 python scripts/bench_kernels.py > gpurun_out/kernels.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/kernels.json'))
for size,res in d['hbm_kernels'].items():
    print(size, {k.split(' ')[0]: round(v['GB/s']) for k,v in res.items()})
PY
