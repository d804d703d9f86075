#!/bin/bash
# LSM pair kernel: parity tests, then v1 / v2 timings at the N = 1 and N = 8 per-rank shapes, then the phase timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lsm.py -q -x > gpurun_out/r2b_lsm_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_lsm_tests.log
tail -5 gpurun_out/r2b_lsm_tests.log
{
for v in 0 1; do
  for shape in "32 32" "64 32" "128 32" "256 32"; do
    echo -n "V2=$v  "; LOCOV_B200_LSM_V2=$v timeout 120 python scripts/lsm_probe.py $shape 2>&1 | tail -1
  done
done
echo -n "V2=1 IPT=1  "; LOCOV_B200_LSM_IPT=1 timeout 120 python scripts/lsm_probe.py 32 32 2>&1 | tail -1
echo -n "V2=1 IPT=1  "; LOCOV_B200_LSM_IPT=1 timeout 120 python scripts/lsm_probe.py 256 32 2>&1 | tail -1
for dbg in 3 4; do
echo -n "V2=1 DEBUG=$dbg  "; LOCOV_B200_DEBUG=$dbg timeout 120 python scripts/lsm_probe.py 32 32 2>&1 | tail -1
echo -n "V2=1 DEBUG=$dbg  "; LOCOV_B200_DEBUG=$dbg timeout 120 python scripts/lsm_probe.py 256 32 2>&1 | tail -1
done
} > gpurun_out/r2b_lsm_times.txt 2>&1
cat gpurun_out/r2b_lsm_times.txt
LOCOV_B200_TIMELINE=1 timeout 120 python scripts/gemm_timeline.py --lsm-only > gpurun_out/r2b_lsm_timeline.txt 2>&1
cat gpurun_out/r2b_lsm_timeline.txt
