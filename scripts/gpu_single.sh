#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n1_full.json | cut -c1-400
timeout 300 python bench.py --size 4096 --grid 1 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-int8-split 2>&1 | tail -1 | tee gpurun_out/bench_cfg2.json | cut -c1-300
timeout 600 python bench.py --size 65536 --grid 4 --dtype bf16 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_n1.json | cut -c1-600
timeout 600 python bench.py --workload tallskinny --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg4_n1.json | cut -c1-600
# ncu: int8-split kernel and bf16 kernel on small problems (short replays)
cat > /tmp/ncu_i8.py <<'PY'
import sys; sys.path.insert(0, '.')
import marlin_b200 as mb
from marlin_b200 import _native as nat
rt = mb.Runtime.get()
A = mb.MTUtils.randomBlockMatrix(None, 4096, 4096, 1, 1, seed=1).blocks[0][1]
B = mb.MTUtils.randomBlockMatrix(None, 4096, 4096, 1, 1, seed=2).blocks[0][1]
rt.set_fp64_mode("int8x8", 5)
for _ in range(2): C = A.multiply(B)
rt.set_fp64_mode("native")
Ab, Bb = A.copy(nat.MB_BF16), B.copy(nat.MB_BF16)
for _ in range(2): D = Ab.multiply(Bb)
import torch; torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ozaki_i8|bf16_tcgen05" -c 4 -o gpurun_out/prof_tc -f python /tmp/ncu_i8.py > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
