#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py -x -q 2>&1 | tail -4
echo "--- 2cta"; timeout 300 python scripts/bench_ozaki.py 4096 8192 16384 2>&1 | tail -3 | cut -c1-900
cp gpurun_out/ozaki.json gpurun_out/ozaki_2cta.json
echo "--- 1cta"; MARLIN_B200_TC_2CTA=0 timeout 300 python scripts/bench_ozaki.py 8192 2>&1 | tail -1 | cut -c1-900
cp gpurun_out/ozaki.json gpurun_out/ozaki_1cta.json
timeout 300 python scripts/bench_kernels.py > gpurun_out/kernels.log 2>&1; grep -A1 "fill_uniform" gpurun_out/kernels.log | grep "GB/s"
