#!/bin/bash
# host-pipeline chunk sweep of ca_process at the headline batch with the 16 x 16 row kernels (growth 8 and 4)
export CA_TIERS=1 CA_NOPROFILE=1
for g in 8 4; do for c in 2 3 4; do
  echo "== growth=$g chunks=$c"; CA_TIER_GROWTH=$g CA_IO_CHUNKS=$c timeout 600 python tools/probe.py 16128 192 2>&1 | tail -1 | cut -c1-120
done; done
