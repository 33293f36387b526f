#!/bin/bash
# the rest of the GPU suite on the final library (tools/gpu_r2_final3.sh ran the engine / mirror / live files)
mkdir -p gpurun_out
SECONDS=0
timeout 170 python -m pytest tests/test_tiers_gpu.py tests/test_fullsize_parity_gpu.py tests/test_group_gpu.py tests/test_multigpu_gpu.py -x -q --durations=8 > gpurun_out/r2j_pytest.log 2>&1; echo "tests rc=$? wall ${SECONDS}s"; tail -14 gpurun_out/r2j_pytest.log
