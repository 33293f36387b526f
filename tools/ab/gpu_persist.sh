#!/bin/bash
# persistent MAC schedule: parity, then A/B against one CTA per item
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_tiers_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | grep -v "^.\[3" | tail -8
for p in 0 auto; do
  if [ $p = auto ]; then unset CA_MAC_PERSIST; else export CA_MAC_PERSIST=$p; fi
  CA_TIERS=1 timeout 120 python tools/probe.py 4096 100 2>&1 | tail -2
  CA_TIERS=1 CA_NOPROFILE=1 timeout 120 python tools/probe.py 4096 200 2>&1 | tail -1
  timeout 120 python tools/probe.py 1024 50 2>&1 | tail -1
done
unset CA_MAC_PERSIST
for v in 1 4 6; do CA_MAC_VARIANT=$v CA_TIERS=1 timeout 120 python tools/probe.py 4096 100 2>&1 | tail -1; done
