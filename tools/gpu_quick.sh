#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^.\[3" | tail -4
export CA_TIERS=1
timeout 120 python tools/probe.py 4096 100 2>&1 | tail -2
for v in 6 7; do CA_MAC_VARIANT=$v timeout 120 python tools/probe.py 4096 100 2>&1 | tail -1; done
CA_NOPROFILE=1 timeout 120 python tools/probe.py 4096 200 2>&1 | tail -1
CA_NOPROFILE=1 CA_MAC_VARIANT=6 timeout 120 python tools/probe.py 4096 200 2>&1 | tail -1
