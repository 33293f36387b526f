#!/bin/bash
export CA_TIERS=1
for g in 4 8; do for v in 13 15 16 17; do
  echo "== growth=$g variant=$v profile"; CA_TIER_GROWTH=$g CA_MAC_VARIANT=$v timeout 600 python tools/probe.py 16128 64 2>&1 | tail -2 | head -1 | cut -c60-330
done; done
