#!/bin/bash
# Multi-GPU visit: NCCL parity test + bench at N = all visible GPUs (and N=1 on the same box for the scaling ratio).
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
for n in $N 1; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_multi_n1.json | cut -c1-600
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 2>&1 | grep -E '^\{|Error|error' | tail -3 | tee gpurun_out/bench_multi_n$n.json | cut -c1-3000
  fi
done
