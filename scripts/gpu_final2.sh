#!/bin/bash
# last single-GPU visit of the round: full regression, then ncu of the two rewritten HBM kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 60 scripts/bin/dms_suite | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:transpose_f64_tile|fill_uniform" -c 6 \
   -f -o gpurun_out/hbm_v3_kernels python scripts/ncu_hbm.py > gpurun_out/ncu_hbm_v3.log 2>&1; tail -2 gpurun_out/ncu_hbm_v3.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600
