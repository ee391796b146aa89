#!/bin/bash
# round-2 one-GPU check: torch-free dist engine with 2 ranks sharing the GPU, then the full default bench line
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=60
(timeout 700 python -m pytest tests/test_gpu_dist_cabi.py -x -q 2>&1 | tail -40) > gpurun_out/r02_dist_cabi_1gpu.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
tail -25 gpurun_out/r02_dist_cabi_1gpu.log
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n1_b.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity']['max_scaled_err'])
        for k, v in (d.get('extra_configs') or {}).items():
            print(k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'steps', 'error')}, (v.get('parity') or {}).get('max_scaled_err'), (v.get('roofline') or {}).get('frac'), (v.get('clocks') or {}).get('sm_mhz'))
PY
tail -5 gpurun_out/r02_bench_n1_b.err
