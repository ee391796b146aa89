#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -30
timeout 600 python scripts/bench_kernels.py > gpurun_out/kernels.log 2>&1; tail -5 gpurun_out/kernels.log
