#!/bin/bash
export CA_TIERS=1 CA_NOPROFILE=1
for c in 3 4; do echo "== chunks=$c"; CA_IO_CHUNKS=$c CA_IO_TRACE=1000 timeout 600 python tools/probe.py 16128 192 2>&1 | grep -E "io trace|K=" | cut -c1-160; done
