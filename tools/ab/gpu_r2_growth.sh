#!/bin/bash
# tier growth 8 (3 tiers, 319 KB / instance-period) vs growth 4 (4 tiers, 258 KB) with the 16 x 16 row kernels
export CA_TIERS=1
for K in 4096 16128; do
for g in 8 4; do
  echo "== K=$K growth=$g profile"; CA_TIER_GROWTH=$g timeout 600 python tools/probe.py $K 64 2>&1 | tail -2 | head -1 | cut -c1-330
  echo "== K=$K growth=$g noprofile"; CA_TIER_GROWTH=$g CA_NOPROFILE=1 timeout 600 python tools/probe.py $K 128 2>&1 | tail -1 | cut -c1-200
done; done
