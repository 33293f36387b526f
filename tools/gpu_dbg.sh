#!/bin/bash
mkdir -p gpurun_out
( time python -c "import torch; print(torch.cuda.is_available())" ) > gpurun_out/dbg_import.log 2>&1
timeout 200 python -u -m pytest tests/test_tiers_gpu.py -m gpu -x -q -s -k persistent -o faulthandler_timeout=40 > gpurun_out/dbg.log 2>&1
echo "pytest rc=$?" >> gpurun_out/dbg.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader >> gpurun_out/dbg.log 2>&1
sort gpurun_out/dbg.log | uniq -c | sort -rn | head -30
