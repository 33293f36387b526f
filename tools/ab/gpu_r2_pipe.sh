#!/bin/bash
# two-lane pipelined schedule with the row-FFT kernels: MAC staging variants x CTAs per SM
mkdir -p gpurun_out
export CA_TIERS=1 CA_NOPROFILE=1
run() { echo "== $*"; env "$@" timeout 300 python tools/probe.py 4096 128 2>&1 | tail -1 | cut -c1-330; }
run CA_PIPELINE=0
run CA_PIPELINE=1
run CA_PIPELINE=1 CA_MAC_VARIANT=5
run CA_PIPELINE=1 CA_MAC_VARIANT=0
run CA_PIPELINE=1 CA_MAC_VARIANT=2
run CA_PIPELINE=1 CA_MAC_VARIANT=1
run CA_PIPELINE=1 CA_MAC_CTAS=3
run CA_PIPELINE=1 CA_MAC_CTAS=2
run CA_PIPELINE=0 CA_MAC_VARIANT=5
echo "== trace variant 5"; CA_PIPELINE=1 CA_MAC_VARIANT=5 CA_PIPE_TRACE=800 timeout 300 python tools/probe.py 4096 64 2>&1 | grep -E "trace|K=" | cut -c1-200
