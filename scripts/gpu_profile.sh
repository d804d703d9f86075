#!/bin/bash
# ncu evidence for profiles/: launch lists (gpu__time_duration) and full captures of the hot kernels.  One GPU.
# usage (GPU box): bash scripts/gpu_profile.sh <tag>     -> gpurun_out/<tag>_*
TAG=${1:-r1c}
O=gpurun_out
mkdir -p $O
# 1. launch list of the headline bench (same command as the bench line, few steps)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/${TAG}_launches_bench.log 2>&1
# 2. full capture of the LSM step kernels (graph replays of the timed region: skip the eager warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:tc_gemm|pair_ce|lsm_prep' -s 4 -c 12 \
    -o $O/${TAG}_bench -f python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/${TAG}_ncu_bench.log 2>&1
# 3. launch list of the per-kernel bench (RoIAlign + box predictor chain, configs 1/3/5)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_kernels.csv \
    python scripts/bench_kernels.py --quick > $O/${TAG}_launches_kernels.log 2>&1
# 4. full captures: RoIAlign 14x14 (config 1) and 7x7 (BASELINE-literal), box scoring at config 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_v4 -s 16 -c 1 -o $O/${TAG}_roi14 -f \
    python scripts/bench_kernels.py --roi-only > $O/${TAG}_ncu_roi14.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_v4 -s 42 -c 1 -o $O/${TAG}_roi7 -f \
    python scripts/bench_kernels.py --roi-only > $O/${TAG}_ncu_roi7.log 2>&1
BK_ONLY=cfg5-bf16 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:tc_gemm|box_score' -s 9 -c 4 -o $O/${TAG}_box5 -f \
    python scripts/bench_kernels.py --box-only > $O/${TAG}_ncu_box5.log 2>&1
ls -la $O | grep ${TAG}
