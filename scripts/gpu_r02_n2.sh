#!/bin/bash
# round-2 two-GPU check: C-ABI dist engine (torch-free), python multi-GPU worker, grouped-kernel regression, bench N=1/2
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=60
(timeout 600 python -m pytest tests/test_gpu_dist_cabi.py -x -q 2>&1 | tail -40) > gpurun_out/r02_n2_dist_cabi.log
(timeout 300 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_matrix_api.py -x -q 2>&1 | tail -15) > gpurun_out/r02_n2_cabi.log
(timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -40) > gpurun_out/r02_n2_multi.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-int8-split --no-e2e > gpurun_out/r02_n2_bench_n1.json 2> gpurun_out/r02_n2_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_n2_bench_n2.json 2> gpurun_out/r02_n2_bench_n2.err
for f in r02_n2_dist_cabi.log r02_n2_cabi.log r02_n2_multi.log; do echo "== $f"; tail -12 gpurun_out/$f; done
echo "== n1"; head -c 400 gpurun_out/r02_n2_bench_n1.json; tail -3 gpurun_out/r02_n2_bench_n1.err
echo "== n2"; head -c 400 gpurun_out/r02_n2_bench_n2.json; tail -5 gpurun_out/r02_n2_bench_n2.err
