#!/bin/bash
# full GPU pass: parity suite, smoke, bench (N=1)
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --durations=8) > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/r2f_pytest_gpu.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_smoke.log; tail -2 gpurun_out/r2f_smoke.log
(time timeout 900 python bench.py) > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -3 gpurun_out/r2f_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2f_bench_n1.json').read().strip().split('\n')[0])
    print('headline', d['ms_per_step'], d['value'], d['dtype'])
    for k,v in d['precisions'].items(): print(' ', k, v['ms_per_step'], v['kernels_ms'])
    print('roofline', d['roofline']['kernel'], d['roofline']['frac'])
    print('e2e', d['e2e']['ms_per_step'])
    for k,v in d['workloads'].items(): print(' ', k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','frac','achieved','fp32','bf16')})
    print('cpu', d['cpu_baseline'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2f_bench_n1.json').read()[:2000])
PY
