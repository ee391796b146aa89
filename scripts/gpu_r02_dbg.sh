#!/bin/bash
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=20
B="bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --no-int8-split --no-extra-configs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 $B > gpurun_out/r02_dbg_a.json 2> gpurun_out/r02_dbg_a.err
echo "== default connections"; grep -h "status 0x\|MarlinError" gpurun_out/r02_dbg_a.err | head -4; head -c 300 gpurun_out/r02_dbg_a.json
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 $B > gpurun_out/r02_dbg_b.json 2> gpurun_out/r02_dbg_b.err
echo "== 32 connections"; grep -h "status 0x\|MarlinError" gpurun_out/r02_dbg_b.err | head -4; head -c 300 gpurun_out/r02_dbg_b.json
python - <<'PY'
import json
for f in ('a', 'b'):
    for l in open(f'gpurun_out/r02_dbg_{f}.json'):
        if l.startswith('{'):
            d = json.loads(l); print(f, 'value', d['value'], 'e2e', d['e2e'])
PY
