#!/bin/bash
# A/B of the exchange-A arrangement (side stream + early signal vs in-order), N GPUs
N=${1:-2}
mkdir -p gpurun_out
for side in 1 0; do
  LOCOV_B200_SYMM_SIDE=$side timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_nccl.py bench 2>&1 | grep -E "parity|Error|error|warn" | tail -3
  LOCOV_B200_SYMM_SIDE=$side timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 10 > gpurun_out/r2s_bench_n${N}_side$side.json 2> gpurun_out/r2s_bench_n${N}_side$side.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s_bench_n${N}_side$side.json').read().strip().split('\n')[-1])
    print('N=${N} side=$side', {k: round(v['ms_per_step'],4) for k,v in d['precisions'].items()}, 'launches', d['gpu_launches_per_step'], 'e2e', round(d['e2e']['ms_per_step'],3)); [print('   kernels', pr, {k: round(v,4) for k,v in d['precisions'][pr]['kernels_ms'].items()}) for pr in d['precisions']]; print('   workloads', {k: {p: round(v[p]['ms'],4) for p in ('fp32','bf16')} for k,v in d.get('workloads',{}).items()})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2s_bench_n${N}_side$side.err').read()[-1500:])
PY
done
