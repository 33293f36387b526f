#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^.\[3" | tail -4
export CA_TIERS=1
timeout 120 python tools/probe.py 4096 100 2>&1 | tail -2
CA_NOPROFILE=1 timeout 120 python tools/probe.py 4096 300 2>&1 | tail -1
