#!/bin/bash
# the G = 2 cases of engine.ir_split / ca_group_reset (gpurun --gpus 2)
mkdir -p gpurun_out
SECONDS=0
timeout 120 python -m pytest "tests/test_group_gpu.py::test_group_reset_restarts_every_member_in_step" "tests/test_dropin_gpu.py::test_ir_split_group_through_the_class_api" "tests/test_group_gpu.py::test_group_small_ir_matches_fp64_and_single_gpu" -q -rs -k "not 4 and not 8" > gpurun_out/r2k_pytest_2gpu.log 2>&1; echo "tests rc=$? wall ${SECONDS}s"; tail -12 gpurun_out/r2k_pytest_2gpu.log | cut -c1-220
