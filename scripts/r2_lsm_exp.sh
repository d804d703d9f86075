#!/bin/bash
for e in 3 4 6; do
  echo "EXP=$e timeline"; LOCOV_B200_LIBDIR=lib_exp$e LOCOV_B200_NVCC_EXTRA="-DLOCOV_EXP=$e" LOCOV_B200_TIMELINE=1 timeout 120 python scripts/gemm_timeline.py --lsm-only 2>&1 | grep -E "acc_done|e8|e10|e11|e12|e13|epi_done"
done
