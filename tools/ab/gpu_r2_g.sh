#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2g_pytest.log | cut -c1-300
export CA_TIERS=1
echo "== profile K=4096"; timeout 300 python tools/probe.py 4096 64 2>&1 | tail -2 | head -1 | cut -c60-300
echo "== profile K=16128"; timeout 600 python tools/probe.py 16128 64 2>&1 | tail -2 | head -1 | cut -c60-300
echo "== K=16128"; CA_NOPROFILE=1 timeout 600 python tools/probe.py 16128 192 2>&1 | tail -1 | cut -c1-120
