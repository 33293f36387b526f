#!/bin/bash
# 2 x B200: multi-GPU parity tests, the scaling line the driver computes (N = 2), the IR-split mode
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 500 python -m pytest tests/test_multigpu_gpu.py tests/test_engine_gpu.py -m gpu -x -q -k "multigpu or edge_cases or two_gpu or nccl or irsplit or instances" 2>&1 | grep -v "^.\[3" | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/two_bench.json 2> gpurun_out/two_bench.err; echo "bench2 rc=$?"; tail -c 1200 gpurun_out/two_bench.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode irsplit --steps 200 --warmup 20 > gpurun_out/two_irsplit.json 2> gpurun_out/two_irsplit.err; echo "irsplit rc=$?"; tail -c 900 gpurun_out/two_irsplit.json
