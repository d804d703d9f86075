#!/bin/bash
# final multi-GPU pass of the round: sharded parity at the bench shape, then the N-GPU bench line (peer-memory exchange path)
N=${1:-2}
mkdir -p gpurun_out
LOCOV_B200_SYMM=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_nccl.py bench 2>&1 | grep -E "parity|Error|error|warn" | tail -4
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_sharded_nccl.py -m gpu -q 2>&1 | tail -3; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 10 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}.json').read().strip().split('\n')[-1])
    print('N=${N}', 'ms_per_step', round(d['ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in d['precisions'].items()}, 'launches', d['gpu_launches_per_step'], 'e2e', round(d['e2e']['ms_per_step'],3), 'value', d['value'])
    for pr in d['precisions']: print('   kernels', pr, {k: round(v,4) for k,v in d['precisions'][pr]['kernels_ms'].items()})
    print('   workloads', {k: {p: round(v[p]['ms'],4) for p in ('fp32','bf16')} for k,v in d.get('workloads',{}).items()})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2_bench_n${N}.err').read()[-1500:])
PY
