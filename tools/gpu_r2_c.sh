#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/r2c_pytest.log | cut -c1-400
export CA_TIERS=1
echo "== profile K=4096"; timeout 300 python tools/probe.py 4096 64 2>&1 | tail -2 | cut -c1-400
echo "== noprofile K=4096"; CA_NOPROFILE=1 timeout 300 python tools/probe.py 4096 128 2>&1 | tail -1 | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2c_bench.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2c_bench.json').read().strip().split('\n')[-1])
    for k in ('value','ms_per_step','steps','e2e','parity_check','sustained_channels','clocks','cfg4_1024_streams_2s','e2e_host_ceiling','irsplit_60s','roofline'):
        print(k, json.dumps(j.get(k))[:700])
    print('lat', json.dumps(j.get('latency_1_instance'))[:900])
except Exception as ex: print('parse fail',ex)
PY
