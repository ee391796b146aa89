#!/bin/bash
# round-2 two-GPU validation: torch-free dist engine (device path with the two-launch fused reduce, host path with pushed
# partials; bench-size cases), python multi-GPU worker, bench N=2
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=30
(timeout 900 python -m pytest tests/test_gpu_dist_cabi.py -x -q 2>&1 | tail -40) > gpurun_out/r02_n2_dist_cabi.log
tail -12 gpurun_out/r02_n2_dist_cabi.log
(timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -20) > gpurun_out/r02_n2_multi.log
tail -4 gpurun_out/r02_n2_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-extra-configs --no-int8-split > gpurun_out/r02_n2_bench_n2.json 2> gpurun_out/r02_n2_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_n2_bench_n2.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity']['max_scaled_err'])
        print('e2e', d['e2e'])
PY
tail -3 gpurun_out/r02_n2_bench_n2.err | cut -c1-200
