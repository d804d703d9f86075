#!/bin/bash
N=${1:-2}
for dbg in 0 1 2; do
  LOCOV_B200_SYMM_DEBUG=$dbg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 20 --no-workloads --precision bf16 > gpurun_out/r2h_dbg$dbg.json 2> gpurun_out/r2h_dbg$dbg.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2h_dbg$dbg.json').read().strip().split('\n')[-1])
    print('N=${N} dbg=$dbg', 'ms_per_step', {k: round(v['ms_per_step'],4) for k,v in d['precisions'].items()}, 'kernels', {k: round(v,4) for k,v in d['precisions']['bf16']['kernels_ms'].items()})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2h_dbg$dbg.err').read()[-800:])
PY
done
