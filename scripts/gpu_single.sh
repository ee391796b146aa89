#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_cabi.py -x -q -k "transpose or elementwise or sum" 2>&1 | tail -2
for v in 0 1; do MARLIN_B200_TRANSPOSE_VARIANT=$v timeout 200 python scripts/bench_kernels.py > gpurun_out/kernels_v$v.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/kernels.json'))
for size,res in d['hbm_kernels'].items():
    print(size, {k.split(' ')[0]: round(v['GB/s']) for k,v in res.items() if k.startswith('transpose')})
PY
done
