#!/bin/bash
# engine.ir_split / ca_group_reset: group + mirror suites (on 1 GPU the G = 2 cases skip; run again with gpurun --gpus 2)
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
SECONDS=0
timeout 170 python -m pytest tests/test_group_gpu.py tests/test_dropin_gpu.py -x -q -rs > gpurun_out/r2k_pytest_${N}gpu.log 2>&1; echo "GPUs $N tests rc=$? wall ${SECONDS}s"; tail -14 gpurun_out/r2k_pytest_${N}gpu.log | cut -c1-220
