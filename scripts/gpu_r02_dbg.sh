#!/bin/bash
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=25 CUDA_DEVICE_MAX_CONNECTIONS=32 MB_BIG_FUSED=8192 MB_SKIP_SMALL=1
S=$(date +%s)
for variant in v2a v2b v2c v3a v3b v3c v1a; do
  export MARLIN_B200_FUSED_SPLIT=${variant:1:1}
  for r in 0 1; do timeout 150 python tests/dist_cabi_worker.py $r 2 dbg${variant}$S 2 > gpurun_out/r02_dbgf_${variant}_$r.log 2>&1 & done
  wait
  echo "== $variant: $(grep -h 'big fused' gpurun_out/r02_dbgf_${variant}_0.log gpurun_out/r02_dbgf_${variant}_1.log | awk '{print $NF}' | tr '\n' ' ')"
  grep -h "diagnosis\|tiles touched\|MarlinError" gpurun_out/r02_dbgf_${variant}_0.log gpurun_out/r02_dbgf_${variant}_1.log | head -3
done
