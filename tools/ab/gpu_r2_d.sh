#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_group_gpu.py tests/test_engine_gpu.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2d_pytest.log | cut -c1-600
for k in 8 32 128; do CA_ENGINE_SHARED=$k CA_ENGINE_TIERS=auto CA_ENGINE_PERIOD=256 timeout 300 python bench.py --mode class-api --class-k $k --class-periods 500 2>&1 | tail -2 | cut -c1-400; done
