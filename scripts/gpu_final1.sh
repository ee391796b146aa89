#!/bin/bash
# single-GPU regression of the round: tests, C++ suite, smoke, transpose variants, kernel bandwidths, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 60 scripts/bin/dms_suite | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python scripts/bench_transpose.py 16 17 18 19 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l and l[0].isdigit():
        v,js=l.split(' ',1); d=json.loads(js)
        print(v, {k:(round(x['GB/s']),x['exact']) for k,x in d.items()} if 'error' not in d else d)
    else: print(l)
"
timeout 200 python scripts/bench_kernels.py hbm > gpurun_out/kernels_v3.log 2>&1; cp gpurun_out/kernels.json gpurun_out/kernels_hbm_v3.json
timeout 400 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n1_v3.json | cut -c1-1800
