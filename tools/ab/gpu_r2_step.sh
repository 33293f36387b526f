#!/bin/bash
# quick check: tier tests + headline-batch probe (profile + real schedule)
timeout 600 python -m pytest tests/test_tiers_gpu.py tests/test_fullsize_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
export CA_TIERS=1
echo "== profile"; timeout 600 python tools/probe.py 16128 64 2>&1 | tail -2 | head -1 | cut -c60-330
echo "== noprofile"; CA_NOPROFILE=1 timeout 600 python tools/probe.py 16128 192 2>&1 | tail -1 | cut -c1-120
