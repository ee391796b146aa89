#!/bin/bash
# Multi-GPU visit: NVLink parity test (both transports) + bench at N = all visible GPUs (and N=1 for the scaling ratio).
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
echo "GPUs: $N"
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25
for tr in p2p nccl; do
  MARLIN_B200_TRANSPORT=$tr timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 $EXTRA 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 | tee gpurun_out/bench_multi_n${N}_$tr.json | cut -c1-2600
done
if [ "$SKIP_N1" != "1" ]; then
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_multi_n1.json | cut -c1-2600
fi
