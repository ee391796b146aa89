#!/bin/bash
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
echo "GPUs: $N"
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -6
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -2 | tee gpurun_out/bench_n${N}_$name.json | cut -c1-240; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$name.json').read().strip().splitlines()[-1])
    print('   ->', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms', d.get('phases_ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), d.get('clocks'))
except Exception as e: print('   parse failed', e)
PY
}
run cfg3_p2p --steps 10 --warmup 3
MARLIN_B200_TRANSPORT=nccl run cfg3_nccl --steps 10 --warmup 3 --no-e2e
run cfg5_bf16 --size 65536 --grid 4 --dtype bf16 --steps 5 --warmup 2
run cfg4_tallskinny --workload tallskinny --steps 10 --warmup 3
