#!/bin/bash
# final check of the round: smoke, full GPU suite, default bench line (what the driver runs), reference arm
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
SECONDS=0; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; echo "suite wall ${SECONDS}s"
SECONDS=0; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; echo "bench wall ${SECONDS}s"; grep -v "^\s" gpurun_out/r2f_bench.err | tail -5 | cut -c1-300
SECONDS=0; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_reference.json 2> gpurun_out/r2f_reference.err; echo "reference rc=$? wall ${SECONDS}s"; tail -1 gpurun_out/r2f_reference.json | cut -c1-400
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2f_bench.json').read().strip().split('\n')[-1])
for k in ('value','ms_per_step','steps','sustained_channels','gpu_launches','library','clocks'):
    print(k, json.dumps(j.get(k))[:300])
print('config tiers', j['config'].get('tiers'), j['config'].get('instances_per_gpu'))
print('e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
print('parity', j['parity_check']['rel_l2_max'])
print('roofline', {k:j['roofline'].get(k) for k in ('achieved','frac','traffic','instances','algorithmic_bytes_per_period','kernel_us_per_period','share_of_step','step_us')})
print('lat', {k:(v['p50_us'],v['p99_us']) for k,v in j['latency_1_instance'].items() if isinstance(v,dict) and 'paced' in v})
print('irsplit', j['irsplit_60s'].get('p2p_fused'))
print('class', j.get('dropin_class_api',{}).get('best_rt_channels'))
print('sust', json.dumps(j['sustained_through_ca_process'])[:600])
print('cfg4', json.dumps(j.get('cfg4_1024_streams_2s'))[:400])
PY
