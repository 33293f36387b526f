#!/bin/bash
# round 2, first call: the whole GPU suite (with the new full-size parity tests) + one default bench line
mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; nvidia-smi --query-gpu=name,memory.total --format=csv >> gpurun_out/r2a_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -25 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2a_bench.json').read().strip().split('\n')[-1])
    print({k:j[k] for k in ('value','ms_per_step','steps','e2e','parity_check','sustained_channels','clocks') if k in j})
except Exception as ex: print('parse fail',ex)
PY
