#!/bin/bash
export CA_TIERS=1
for g in 8 4; do for v in 6 7; do
  echo "== growth=$g variant=$v profile"; CA_TIER_GROWTH=$g CA_MAC_VARIANT=$v timeout 600 python tools/probe.py 16128 64 2>&1 | tail -2 | head -1 | cut -c60-330
done; done
for v in 6 7; do echo "== K=4096 growth=8 variant=$v"; CA_TIER_GROWTH=8 CA_MAC_VARIANT=$v timeout 600 python tools/probe.py 4096 64 2>&1 | tail -2 | head -1 | cut -c60-330; done
