#!/bin/bash
mkdir -p gpurun_out
export CA_TIERS=1 CA_NOPROFILE=1
for c in 4 3 2; do
echo "== trace pipeline, MAC CTAs/SM=$c"; CA_PIPELINE=1 CA_MAC_CTAS=$c CA_PIPE_TRACE=800 timeout 300 python tools/probe.py 4096 64 2>&1 | grep -E "trace|K=" | cut -c1-200
done
