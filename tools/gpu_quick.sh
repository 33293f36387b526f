#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^.\[3" | tail -4
export CA_TIERS=1 CA_NOPROFILE=1
for p in 0 1 0 1; do CA_PDL=$p timeout 120 python tools/probe.py 4096 300 2>&1 | tail -1 | sed "s/^/pdl=$p /"; done
