#!/bin/bash
# round-2 eight-GPU run (final code): headline at N=8 with parity, e2e and the extra configs, then the 8-rank torch-free
# C-ABI test and the Python multi-GPU worker over the peer-memory transport.  Every wait is bounded (30 s).
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=30
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "== n8 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n8.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('n8 value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'], 'launches', d['gpu_launches'])
        print('n8 e2e', d['e2e'])
        print('n8 phases', d['phases_ms_per_step'], 'clocks', d['clocks'])
        for k, v in (d.get('extra_configs') or {}).items():
            print('n8', k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'steps', 'error')}, (v.get('parity') or {}).get('max_scaled_err'), (v.get('roofline') or {}).get('frac'))
PY
(timeout 300 python -m pytest tests/test_gpu_dist_cabi.py -x -q -k "all_gpus or fused" 2>&1 | tail -8) > gpurun_out/r02_pytest_dist_cabi_8gpu.log
tail -4 gpurun_out/r02_pytest_dist_cabi_8gpu.log
(timeout 200 python -m pytest tests/test_gpu_multi.py -x -q -k "p2p" 2>&1 | tail -8) > gpurun_out/r02_pytest_multi_p2p_8gpu.log
tail -3 gpurun_out/r02_pytest_multi_p2p_8gpu.log
