#!/bin/bash
# round-2 profiling pass on ONE GPU (numbers printed under ncu are never bench values):
#  1. launch list of the default bench command (kernel shares of the step)
#  2. ncu --set full of the grouped DMMA kernel as the bench launches it (16384^2, 8 products in one launch): DRAM bytes
#  3. ncu --set full of the generator's fast kernel and the transpose at 16384^2
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-int8-split --no-parity --no-extra-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_dmma_grouped -s 1 -c 1 -o gpurun_out/r02_gemm_grouped -f $B > gpurun_out/r02_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fill_uniform_fast|transpose_f64_tile" -c 4 -o gpurun_out/r02_hbm -f python scripts/bench_kernels.py > gpurun_out/r02_ncu_hbm.log 2>&1
timeout 300 python scripts/bench_kernels.py > gpurun_out/r02_kernels_hbm.json 2> gpurun_out/r02_kernels_hbm.err
for r in r02_gemm_grouped r02_hbm; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv 2>/dev/null | python - "$r" <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
if not rows: sys.exit()
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[1:]:
    print(sys.argv[1], {hdr[i]: r[i] for i in idx})
PY
done | tee gpurun_out/r02_ncu_summary.txt
tail -3 gpurun_out/r02_ncu_gemm.log; head -c 1500 gpurun_out/r02_kernels_hbm.json
