#!/bin/bash
# final check of the round: smoke, full GPU suite, default bench line (what the driver runs), reference arm
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
echo skip-suite
SECONDS=0; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; echo "bench wall ${SECONDS}s"; grep -v "^\s" gpurun_out/r2z_bench.err | tail -5 | cut -c1-300
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2z_bench.json').read().strip().split('\n')[-1])
for k in ('value','ms_per_step','steps','sustained_channels','gpu_launches','library','clocks'):
    print(k, json.dumps(j.get(k))[:300])
print('e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
print('parity', j['parity_check']['rel_l2_max'])
print('roofline', {k:j['roofline'][k] for k in ('achieved','frac','traffic','instances','algorithmic_bytes_per_period','kernel_us_per_period','share_of_step')})
print('lat', {k:(v['p50_us'],v['p99_us']) for k,v in j['latency_1_instance'].items() if isinstance(v,dict) and 'paced' in v})
print('irsplit', j['irsplit_60s'].get('p2p_fused'))
print('class', j.get('dropin_class_api',{}).get('best_rt_channels'))
print('sust', json.dumps(j['sustained_through_ca_process'])[:500])
PY
