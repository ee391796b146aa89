#!/bin/bash
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=25
echo "== fence_all=0 (old behaviour)"
(MARLIN_B200_FENCE_ALL=0 timeout 300 python -m pytest tests/test_gpu_dist_cabi.py -x -q -k "fused_reduce" 2>&1 | grep -E "big fused|passed|failed|Error" | head -12)
echo "== fence_all=1"
(timeout 300 python -m pytest tests/test_gpu_dist_cabi.py -x -q -k "fused_reduce" 2>&1 | grep -E "big fused|passed|failed|Error" | head -12)
