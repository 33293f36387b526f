#!/bin/bash
# round-2 final evidence on one GPU: smoke, GPU suite, default bench line + reference arm, ncu launch list of one tier
# cycle at the headline batch, ncu --set full of the MAC launches and of the FFT kernels of one period (K = 4096)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
SECONDS=0; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; echo "suite wall ${SECONDS}s"
SECONDS=0; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$? wall ${SECONDS}s"; grep -v "^\s" gpurun_out/r2f_bench.err | tail -3 | cut -c1-300
SECONDS=0; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_reference.json 2> gpurun_out/r2f_reference.err; echo "reference rc=$? wall ${SECONDS}s"
CMD="python bench.py --steps 64 --warmup 3 --no-latency --no-sustained --no-cpu-baseline --no-parity --no-cfg4 --no-host-ceiling --no-irsplit --no-class-api --no-roofline"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fwd0|k_inv0|k_mac|k_tfwd|k_tinv|k_tcols|k_trows' -s 9097 -c 704 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_launches.log 2>&1; echo "launch list rc=$?"
export CA_TIERS=1
P="python tools/probe.py 4096 8"
timeout 200 $P > gpurun_out/r2f_probe_plain.log 2>&1; echo "probe rc=$?"; tail -2 gpurun_out/r2f_probe_plain.log | head -1 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mac_p' -s 2280 -c 3 -o gpurun_out/r2f_prof_mac $P > gpurun_out/r2f_ncu_mac.log 2>&1; echo "mac rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'x2|k_tcols' -s 6080 -c 8 -o gpurun_out/r2f_prof_fft $P > gpurun_out/r2f_ncu_fft.log 2>&1; echo "fft rc=$?"
ls -la gpurun_out/r2f_prof_*.ncu-rep
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2f_bench.json').read().strip().split('\n')[-1])
for k in ('value','ms_per_step','steps','sustained_channels','gpu_launches','clocks'):
    print(k, json.dumps(j.get(k))[:300])
print('e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
print('parity', j['parity_check']['rel_l2_max'])
print('roofline', {k:j['roofline'].get(k) for k in ('achieved','frac','traffic','instances','algorithmic_bytes_per_period','kernel_us_per_period','share_of_step','step_us')})
print('lat', {k:(v['p50_us'],v['p99_us']) for k,v in j['latency_1_instance'].items() if isinstance(v,dict) and 'paced' in v})
print('irsplit', j['irsplit_60s'].get('p2p_fused'))
print('class', j.get('dropin_class_api',{}).get('best_rt_channels'))
print('sust', json.dumps(j['sustained_through_ca_process'])[:500])
r=json.loads(open('gpurun_out/r2f_reference.json').read().strip().split('\n')[-1])
print('reference', r['value'], r.get('clocks'))
PY
