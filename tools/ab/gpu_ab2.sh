#!/bin/bash
# A/B of library builds (CA_B200_LIB): profile at K=4096, device/e2e wall at K=4096 and K=16128, twice each
export CA_TIERS=1
for v in "$@"; do
  echo "== lib$v profile K=4096"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so timeout 300 python tools/probe.py 4096 64 2>&1 | tail -2 | head -1 | cut -c60-300
  echo "== lib$v K=4096"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so CA_NOPROFILE=1 timeout 300 python tools/probe.py 4096 128 2>&1 | tail -1 | cut -c1-120
done
for r in 1 2; do for v in "$@"; do
  echo "== lib$v K=16128 run $r"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so CA_NOPROFILE=1 timeout 600 python tools/probe.py 16128 192 2>&1 | tail -1 | cut -c1-120
done; done
