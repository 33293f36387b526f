#!/bin/bash
# host-mirror changes (engine.shared rendezvous, onStop/leave, IR lock): the GPU tests that go through the mirror + the class-API bench leg
mkdir -p gpurun_out
SECONDS=0
timeout 230 python -m pytest tests/test_dropin_gpu.py tests/test_live_gpu.py -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "tests rc=$? wall ${SECONDS}s"; tail -4 gpurun_out/r2h_pytest.log
SECONDS=0
CA_ENGINE_SHARED=32 CA_ENGINE_TIERS=auto CA_ENGINE_PERIOD=256 timeout 90 python bench.py --mode class-api --class-k 32 --class-periods 1000 > gpurun_out/r2h_class32.json 2> gpurun_out/r2h_class32.err; echo "class-api rc=$? wall ${SECONDS}s"; tail -1 gpurun_out/r2h_class32.json
