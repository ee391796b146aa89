#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_n1_final.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1_final.json').read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'roof', d['roofline']['frac'], d['roofline']['traffic'], 'int8', d['fp64_on_int8_tensor_cores']['value'], d['fp64_on_int8_tensor_cores']['max_err_vs_native_scaled_by_absA_absB'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-int8-split > gpurun_out/ncu_bench_final.log 2>&1; tail -1 gpurun_out/ncu_bench_final.log | cut -c1-200
