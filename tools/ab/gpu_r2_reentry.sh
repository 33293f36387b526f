#!/bin/bash
# re-entry check: smoke, full GPU suite, default bench line
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
SECONDS=0; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; echo "suite wall ${SECONDS}s"
SECONDS=0; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2re_bench.json 2> gpurun_out/r2re_bench.err; echo "bench rc=$?"; echo "bench wall ${SECONDS}s"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2re_bench.json').read().strip().split('\n')[-1])
for k in ('value','ms_per_step','steps','sustained_channels','gpu_launches'):
    print(k, json.dumps(j.get(k))[:300])
print('e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
print('parity', j['parity_check']['rel_l2_max'])
print('roofline', {k:j['roofline'][k] for k in ('achieved','frac','traffic','instances','kernel_us_per_period','share_of_step')})
PY
