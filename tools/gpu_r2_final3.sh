#!/bin/bash
# after the host-side changes of the last session (param_queue.h, ParamMirror, shared rendezvous): smoke + the engine and mirror suites
mkdir -p gpurun_out
SECONDS=0
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2; echo "smoke rc=$? wall ${SECONDS}s"
SECONDS=0
timeout 200 python -m pytest tests/test_engine_gpu.py tests/test_dropin_gpu.py tests/test_live_gpu.py -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "tests rc=$? wall ${SECONDS}s"; tail -4 gpurun_out/r2i_pytest.log
