#!/bin/bash
# A/B of library builds under gpurun_tmp/ (CA_B200_LIB override): per-kernel profile + device/e2e wall at K = 4096
export CA_TIERS=1
for v in "$@"; do
  echo "== lib$v profile"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so timeout 300 python tools/probe.py 4096 64 2>&1 | tail -2 | head -1 | cut -c1-300
  echo "== lib$v noprofile"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so CA_NOPROFILE=1 timeout 300 python tools/probe.py 4096 128 2>&1 | tail -1 | cut -c1-200
done
