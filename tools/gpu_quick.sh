#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^.\[3" | tail -6
timeout 200 python tools/latency.py 256 192000 3000 2>&1 | grep "uniform \|g8 max16384"
timeout 200 python tools/latency.py 64 480000 3000 2>&1 | grep "uniform \|g8 max16384"
