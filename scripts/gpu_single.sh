#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python scripts/bench_kernels.py > gpurun_out/kernels.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/kernels.json'))
for size,res in d['hbm_kernels'].items():
    print(size, {k.split(' ')[0]: round(v['GB/s']) for k,v in res.items()})
PY
timeout 300 ncu --set full --clock-control none -k regex:"binary_flat|unary_flat|transpose_f64|sum_strided" -s 5 -c 5 -o gpurun_out/prof_hbm -f python scripts/ncu_hbm.py > gpurun_out/ncu_hbm.log 2>&1; tail -1 gpurun_out/ncu_hbm.log
