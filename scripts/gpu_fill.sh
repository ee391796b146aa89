#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_matrix_api.py tests/test_gpu_vector.py -x -q 2>&1 | tail -3
timeout 200 python scripts/bench_kernels.py hbm > gpurun_out/kernels_v4.log 2>&1; cp gpurun_out/kernels.json gpurun_out/kernels_hbm_v4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/kernels_hbm_v4.json'))
for k,v in d['hbm_kernels'].items():
    for n,r in v.items():
        if 'fill' in n or 'transpose' in n or 'sum' in n: print(k, n, round(r['GB/s']))
PY
