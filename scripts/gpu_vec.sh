#!/bin/bash
# one GPU visit for the vector kernels: parity, bandwidth, ncu
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_vector.py -x -q 2>&1 | tail -6
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python scripts/bench_kernels.py hbm > gpurun_out/kernels_vec.log 2>&1; tail -3 gpurun_out/kernels_vec.log
cp gpurun_out/kernels.json gpurun_out/kernels_hbm_vec.json 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:gemv|ger_kernel|dot_stage1|fold_partials" -c 12 \
   -f -o gpurun_out/vec_kernels python scripts/ncu_hbm.py > gpurun_out/ncu_vec.log 2>&1; tail -2 gpurun_out/ncu_vec.log
timeout 60 scripts/bin/dms_suite | tail -8
ls -la gpurun_out | tail -5
