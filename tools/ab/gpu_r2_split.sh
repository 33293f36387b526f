#!/bin/bash
# MAC lane and FFT lanes on disjoint SM sets (green contexts) under the two-lane pipelined schedule
export CA_TIERS=1 CA_NOPROFILE=1
run() { echo "== $*"; env "$@" timeout 600 python tools/probe.py ${K:-4096} 128 2>&1 | tail -2 | cut -c1-230; }
K=4096 run CA_PIPELINE=0
for s in 96 80 64 112; do K=4096 run CA_SM_SPLIT=$s; done
K=16128 run CA_PIPELINE=0
for s in 96 80; do K=16128 run CA_SM_SPLIT=$s; done
echo "== trace split 80"; CA_SM_SPLIT=80 CA_PIPE_TRACE=800 timeout 300 python tools/probe.py 4096 64 2>&1 | grep -E "trace|K=" | tail -20 | cut -c1-160
