#!/bin/bash
# pipelined batch schedule: parity, then A/B (device-resident and e2e) at K = 4096
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^.\[3" | tail -8
export CA_TIERS=1 CA_NOPROFILE=1
CA_PIPELINE=0 timeout 120 python tools/probe.py 4096 200 2>&1 | tail -1
CA_PIPELINE=1 timeout 120 python tools/probe.py 4096 200 2>&1 | tail -1
CA_PIPELINE=1 CA_MAC_CTAS=3 timeout 120 python tools/probe.py 4096 200 2>&1 | tail -1
CA_PIPELINE=1 CA_MAC_CTAS=2 timeout 120 python tools/probe.py 4096 200 2>&1 | tail -1
CA_PIPELINE=1 CA_MAC_PERSIST=0 timeout 120 python tools/probe.py 4096 200 2>&1 | tail -1
CA_PIPELINE=1 CA_MAC_CTAS=3 CA_MAC_VARIANT=6 timeout 120 python tools/probe.py 4096 200 2>&1 | tail -1
