#!/bin/bash
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=25
(timeout 400 python -m pytest tests/test_gpu_dist_cabi.py -x -q -k "bench_size" 2>&1 | tail -60) > gpurun_out/r02_dbg_big.log
grep -n "marlin_b200 rank\|big e2e\|passed\|failed" gpurun_out/r02_dbg_big.log | head -40
B="bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --no-int8-split --no-extra-configs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 $B > gpurun_out/r02_dbg_a.json 2> gpurun_out/r02_dbg_a.err
grep -h "marlin_b200 rank\|MarlinError" gpurun_out/r02_dbg_a.err | head -30
python - <<'PY'
import json
for l in open('gpurun_out/r02_dbg_a.json'):
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
PY
