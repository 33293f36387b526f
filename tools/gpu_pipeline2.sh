#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tiers_gpu.py -m gpu -x -q -k "pipelined or persistent" 2>&1 | grep -v "^.\[3" | tail -5
export CA_TIERS=1 CA_NOPROFILE=1 CA_PIPELINE=1 CA_PIPE_TRACE=900
timeout 120 python tools/probe.py 4096 200 2>&1 | grep "trace\|K=" 
CA_MAC_CTAS=3 timeout 120 python tools/probe.py 4096 200 2>&1 | grep "trace\|K=" | tail -12
CA_MAC_CTAS=2 timeout 120 python tools/probe.py 4096 200 2>&1 | grep "K="
CA_PIPE_PRIO=0 timeout 120 python tools/probe.py 4096 200 2>&1 | grep "K="
CA_MAC_PERSIST=0 timeout 120 python tools/probe.py 4096 200 2>&1 | grep "K="
