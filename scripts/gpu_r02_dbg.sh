#!/bin/bash
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=25 CUDA_DEVICE_MAX_CONNECTIONS=32 MB_BIG_FUSED=8192
S=$(date +%s)
for variant in "fence1" "fence0" "fence1b" "slow"; do
  export MARLIN_B200_FENCE_ALL=1 MARLIN_B200_DIST_SLOW=0
  [ $variant = fence0 ] && export MARLIN_B200_FENCE_ALL=0
  [ $variant = slow ] && export MARLIN_B200_DIST_SLOW=1
  for r in 0 1; do MB_SKIP_SMALL=1 timeout 200 python tests/dist_cabi_worker.py $r 2 dbg${variant}$S 2 > gpurun_out/r02_dbgf_${variant}_$r.log 2>&1 & done
  wait
  echo "== $variant"; grep -h "big fused\|diagnosis\|    (\|tiles touched\|Error\|ok worst" gpurun_out/r02_dbgf_${variant}_0.log gpurun_out/r02_dbgf_${variant}_1.log | head -30
done
