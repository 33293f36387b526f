#!/bin/bash
# 16 x 16 row-FFT family: GPU suite, then A/B against the 8 x 8 x 4 rows (CA_ROWS16=0) and register-bound variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
export CA_TIERS=1
echo "== rows8 profile"; CA_ROWS16=0 timeout 300 python tools/probe.py 4096 64 2>&1 | tail -2 | head -1 | cut -c1-300
echo "== rows8 noprofile"; CA_ROWS16=0 CA_NOPROFILE=1 timeout 300 python tools/probe.py 4096 128 2>&1 | tail -1 | cut -c1-200
for v in "$@"; do
  echo "== lib$v profile"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so timeout 300 python tools/probe.py 4096 64 2>&1 | tail -2 | head -1 | cut -c1-300
  echo "== lib$v noprofile"; CA_B200_LIB=$PWD/gpurun_tmp/lib$v.so CA_NOPROFILE=1 timeout 300 python tools/probe.py 4096 128 2>&1 | tail -1 | cut -c1-200
done
