#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_cabi.py -x -q -k "fill_uniform or transpose" 2>&1 | tail -4
timeout 200 python -m pytest tests/test_gpu_matrix_api.py tests/test_gpu_vector.py -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_transpose.py 2 10 11 12 13 14 15 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l and l[0].isdigit():
        v,js=l.split(' ',1); d=json.loads(js)
        print(v, {k:(round(x['GB/s']),x['exact']) for k,x in d.items()} if 'error' not in d else d)
    else: print(l)
"
timeout 200 python scripts/bench_kernels.py hbm > gpurun_out/kernels_v2.log 2>&1; cp gpurun_out/kernels.json gpurun_out/kernels_hbm_v2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/kernels_hbm_v2.json'))
for k,v in d['hbm_kernels'].items():
    for n,r in v.items():
        if 'fill' in n or 'gemv' in n or 'transpose' in n: print(k, n, round(r['GB/s']))
PY
