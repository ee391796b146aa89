#!/bin/bash
# round-2 eight-GPU run: headline at N=8 (device-timed value, e2e through mb_matmul_blocked_dist_host, parity, extra
# configs), the multi-GPU test suites, then N=4 on four of the GPUs.  Every wait is bounded (30 s).
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=30
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
timeout 420 bash -c "$(declare -f run); run 8 29531 --steps 10 --warmup 3 --no-cpu-baseline" > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "== n8 rc=$?"; tail -3 gpurun_out/r02_bench_n8.err | cut -c1-300
(timeout 500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dist_cabi.py -x -q 2>&1 | tail -40) > gpurun_out/r02_pytest_multi_8gpu.log
tail -6 gpurun_out/r02_pytest_multi_8gpu.log
timeout 240 bash -c "$(declare -f run); run 4 29532 --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs --no-int8-split" > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
echo "== n4 rc=$?"
python - <<'PY'
import json
for f in ('n8', 'n4'):
    try:
        for l in open(f'gpurun_out/r02_bench_{f}.json'):
            if l.startswith('{'):
                d = json.loads(l)
                print(f, 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 3), 'parity', d['parity']['max_scaled_err'], 'launches', d['gpu_launches'])
                print(f, 'e2e', d['e2e'])
                print(f, 'phases', d['phases_ms_per_step'], 'clocks', d['clocks'])
                for k, v in (d.get('extra_configs') or {}).items():
                    print(f, k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'steps', 'error')}, (v.get('parity') or {}).get('max_scaled_err'), (v.get('roofline') or {}).get('frac'))
    except Exception as exc:
        print(f, 'unreadable', exc)
PY
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r02_topo_8gpu.txt
