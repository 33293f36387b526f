#!/bin/bash
timeout 600 python -m pytest tests/test_tiers_gpu.py tests/test_fullsize_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
export CA_TIERS=1 CA_NOPROFILE=1
for g in 8 4; do for c in 3 4; do
  echo "== growth=$g chunks=$c"; CA_TIER_GROWTH=$g CA_IO_CHUNKS=$c timeout 600 python tools/probe.py 16128 192 2>&1 | tail -1 | cut -c1-120
done; done
