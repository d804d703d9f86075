#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lsm.py -q > gpurun_out/r2c_lsm_tests.log 2>&1; tail -6 gpurun_out/r2c_lsm_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 3 -c 1 -f -o gpurun_out/r2c_lsm_v2 python scripts/lsm_probe.py 32 32 > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
