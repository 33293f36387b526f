#!/bin/bash
# PDL inside CUDA graphs (single-instance latency mode): parity first, then latency A/B
mkdir -p gpurun_out
CA_PDL=1 timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^.\[3" | tail -4
for p in 0 1; do CA_PDL=$p timeout 200 python tools/latency.py 256 192000 3000 2>&1 | grep "uniform \|g8 max16384" | sed "s/^/pdl=$p /"; done
for p in 0 1; do CA_PDL=$p timeout 200 python tools/latency.py 64 480000 3000 2>&1 | grep "uniform \|g8 max16384" | sed "s/^/pdl=$p /"; done
