#!/bin/bash
# MAC staging variants at growth 4 (short row lists per item), K = 16128
export CA_TIERS=1
for v in 6 4 2 7; do
  echo "== variant=$v profile"; CA_MAC_VARIANT=$v timeout 600 python tools/probe.py 16128 64 2>&1 | tail -2 | head -1 | cut -c60-330
  echo "== variant=$v noprofile"; CA_MAC_VARIANT=$v CA_NOPROFILE=1 timeout 600 python tools/probe.py 16128 128 2>&1 | tail -1 | cut -c1-120
done
