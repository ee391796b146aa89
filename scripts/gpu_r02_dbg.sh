#!/bin/bash
mkdir -p gpurun_out
export MARLIN_B200_TIMEOUT_S=25 CUDA_DEVICE_MAX_CONNECTIONS=32 MB_BIG_FUSED=8192 MB_SKIP_SMALL=1
S=$(date +%s)
for variant in base1 base2 split1 split2 split3 poll1 poll2 poll3; do
  unset MARLIN_B200_FUSED_SPLIT MARLIN_B200_CIN_POLL
  case $variant in split*) export MARLIN_B200_FUSED_SPLIT=1;; poll*) export MARLIN_B200_CIN_POLL=1;; esac
  for r in 0 1; do timeout 150 python tests/dist_cabi_worker.py $r 2 dbg${variant}$S 2 > gpurun_out/r02_dbgf_${variant}_$r.log 2>&1 & done
  wait
  echo "== $variant: $(grep -h 'big fused' gpurun_out/r02_dbgf_${variant}_0.log gpurun_out/r02_dbgf_${variant}_1.log | awk '{print $NF}' | tr '\n' ' ')"
  grep -h "diagnosis\|tiles touched\|MarlinError" gpurun_out/r02_dbgf_${variant}_0.log gpurun_out/r02_dbgf_${variant}_1.log | head -4
done
