#!/bin/bash
export CA_TIERS=1 CA_NOPROFILE=1
for c in 3 4 5 6 8 12; do
  echo "== chunks=$c"; CA_IO_CHUNKS=$c timeout 600 python tools/probe.py 16128 192 2>&1 | tail -1 | cut -c1-120
done
