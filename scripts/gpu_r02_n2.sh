#!/bin/bash
# round-2 two-GPU check: C-ABI dist engine (torch-free, device + host paths), python multi-GPU worker, factorizations,
# generator fast path, bench N=2 with the e2e leg through mb_matmul_blocked_dist_host
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=60
(timeout 600 python -m pytest tests/test_gpu_dist_cabi.py -x -q 2>&1 | tail -40) > gpurun_out/r02_n2_dist_cabi.log
(timeout 600 python -m pytest tests/test_gpu_factor.py tests/test_gpu_cabi.py tests/test_gpu_matrix_api.py tests/test_gpu_bf16.py -x -q 2>&1 | tail -30) > gpurun_out/r02_n2_cabi.log
(timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -40) > gpurun_out/r02_n2_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_n2_bench_n2.json 2> gpurun_out/r02_n2_bench_n2.err
for f in r02_n2_dist_cabi.log r02_n2_cabi.log r02_n2_multi.log; do echo "== $f"; tail -15 gpurun_out/$f; done
python - <<'PY'
import json
for l in open('gpurun_out/r02_n2_bench_n2.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity']['max_scaled_err'])
        print('e2e', d['e2e'])
        for k, v in (d.get('extra_configs') or {}).items():
            print(k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'steps', 'error')}, (v.get('parity') or {}).get('max_scaled_err'), (v.get('roofline') or {}).get('frac'))
PY
tail -5 gpurun_out/r02_n2_bench_n2.err
