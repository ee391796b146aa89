#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py -x -q 2>&1 | tail -6
timeout 300 python scripts/bench_ozaki.py 4096 8192 2>&1 | tail -2 | cut -c1-1500
( nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 200 > gpurun_out/clocks_ozaki.csv & echo $! > /tmp/smi.pid )
timeout 300 python scripts/bench_ozaki.py 16384 2>&1 | tail -1 | cut -c1-1500
kill $(cat /tmp/smi.pid); sort gpurun_out/clocks_ozaki.csv | uniq -c | sort -rn | head -8
