#!/bin/bash
# tier-shape sweep with the final kernels: growth x max block, K = 4096, device-resident and e2e
export CA_TIERS=1 CA_NOPROFILE=1
for g in 8 16 4; do for mb in 16384 8192; do
  CA_TIER_GROWTH=$g CA_TIER_MAXBLOCK=$mb timeout 150 python tools/probe.py 4096 256 2>&1 | tail -1 | sed "s/^/growth=$g max=$mb /" | cut -c1-150
done; done
