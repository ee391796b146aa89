#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_cabi.py -x -q -k "rowsharded or pipelined" 2>&1 | tail -4
timeout 500 python bench.py --workload tallskinny --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg4_n1_e2e.json | cut -c1-2200
