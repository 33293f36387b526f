#!/bin/bash
# headline-batch A/B: default bench line per library / family
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-class-api --no-latency --no-sustained --no-cpu-baseline --no-cfg4 --no-host-ceiling --no-irsplit --no-parity > gpurun_out/x2b_$name.json 2> gpurun_out/x2b_$name.err || { echo "$name failed"; tail -3 gpurun_out/x2b_$name.err; }
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.loads(open(f'gpurun_out/x2b_{n}.json').read().strip().split('\n')[-1])
    print(n, 'value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'parity', j.get('parity_check',{}).get('rel_l2_max'), 'step_us', j['roofline']['step_us'])
except Exception as e: print(n, 'parse failed', e)
PY
}
run rows8 CA_ROWS16=0
for v in "$@"; do run $v CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so; done
