#!/bin/bash
# One GPU visit: tests, smoke, bench, ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 2>&1 | tee gpurun_out/bench_n1.json | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_ref.json | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_dmma -s 8 -c 2 -o gpurun_out/prof_gemm_f64 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
