#!/bin/bash
for d in 16 8; do CA_TIER_DIV=$d timeout 300 python bench.py --steps 100 --warmup 10 --no-latency --no-sustained --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('div=$d value',j['value'],j['ms_per_step'],'e2e',j['e2e']['value'],j['e2e']['ms_per_step'])"; done
