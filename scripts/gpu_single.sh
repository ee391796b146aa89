#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q 2>&1 | tail -15
timeout 600 python scripts/bench_ozaki.py 2048 4096 8192 2>&1 | tail -8
timeout 600 python scripts/bench_kernels.py > gpurun_out/kernels.log 2>&1; grep -E "GB/s|fill|copy|axpb|transpose" gpurun_out/kernels.log | head -40
