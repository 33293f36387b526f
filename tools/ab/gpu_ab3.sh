#!/bin/bash
export CA_TIERS=1
for v in "$@"; do
  echo "== lib$v profile K=4096"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so timeout 300 python tools/probe.py 4096 64 2>&1 | tail -2 | head -1 | cut -c60-300
  echo "== lib$v profile K=16128"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so timeout 600 python tools/probe.py 16128 64 2>&1 | tail -2 | head -1 | cut -c60-300
  echo "== lib$v K=16128"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so CA_NOPROFILE=1 timeout 600 python tools/probe.py 16128 192 2>&1 | tail -1 | cut -c1-120
done
