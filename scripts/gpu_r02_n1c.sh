#!/bin/bash
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=40
(timeout 800 python -m pytest tests/test_gpu_dist_cabi.py -x -q 2>&1 | tail -40) > gpurun_out/r02_dist_cabi_1gpu.log
tail -25 gpurun_out/r02_dist_cabi_1gpu.log
(timeout 300 python -m pytest tests/test_gpu_factor.py tests/test_gpu_matrix_api.py tests/test_cpp_host.py -x -q 2>&1 | tail -15) > gpurun_out/r02_factor_1gpu.log
tail -8 gpurun_out/r02_factor_1gpu.log
