#!/bin/bash
# multi-GPU: ca_group parity tests (small IR + 60 s IR, P2P fused and NCCL), IR-split via torch.distributed, and the N-GPU bench line
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 900 python -m pytest tests/test_group_gpu.py tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/r2m_pytest_${N}gpu.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2m_pytest_${N}gpu.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-roofline > gpurun_out/r2m_bench_${N}gpu.json 2> gpurun_out/r2m_bench_${N}gpu.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2m_bench_${N}gpu.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2m_bench_${N}gpu.json').read().strip().split('\n')[-1])
    for k in ('value','ms_per_step','e2e','parity_check','cfg4_1024_streams_2s','e2e_host_ceiling','irsplit_60s'):
        print(k, json.dumps(j.get(k))[:900])
    print('lat', json.dumps(j.get('latency_1_instance',{}).get('per_gpu'))[:600])
except Exception as ex: print('parse fail',ex)
PY
