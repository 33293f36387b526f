#!/bin/bash
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_parity_gpu.py -m gpu -x -q -k "chunked or batch" 2>&1 | tail -3
export CA_TIERS=1 CA_NOPROFILE=1
for d in 0 1; do for c in 3 4; do echo "== dual=$d chunks=$c"; CA_IO_DUAL=$d CA_IO_CHUNKS=$c CA_IO_TRACE=1000 timeout 600 python tools/probe.py 16128 192 2>&1 | grep -E "io trace|K=" | cut -c1-150; done; done
