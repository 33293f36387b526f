#!/bin/bash
# GPU suite, headline bench (short), ncu launch list of one tier cycle at the headline batch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
CMD="python bench.py --steps 64 --warmup 3 --no-latency --no-sustained --no-cpu-baseline --no-parity --no-cfg4 --no-host-ceiling --no-irsplit --no-class-api"
timeout 600 $CMD > gpurun_out/x2c_plain.json 2> gpurun_out/x2c_plain.err; echo "plain rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/x2c_plain.json').read().strip().split('\n')[-1])
print('value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'step_us', j['roofline']['step_us'], 'frac', j['roofline']['frac'])
PY
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fwd0|k_inv0|k_mac|k_tfwd|k_tinv|k_tcols|k_trows' -s 9097 -c 704 --csv --log-file gpurun_out/x2c_launches.csv $CMD --no-roofline > gpurun_out/x2c_ncu.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/x2c_launches.csv')) if len(r)>14 and r[0].isdigit()]
d=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4].split('(')[0]; d[k][0]+=1; d[k][1]+=float(r[14])/1000
tot=sum(v[1] for v in d.values())
for k,v in sorted(d.items(), key=lambda kv:-kv[1][1]): print(f"{k:40s} n={v[0]:4d} total={v[1]:9.1f}us avg={v[1]/v[0]:7.1f} share={v[1]/tot:.3f}")
print('total', tot, 'per period', tot/64)
PY
