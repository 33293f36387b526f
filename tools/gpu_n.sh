#!/bin/bash
# N-GPU scaling line exactly as the driver launches it: bash tools/gpu_n.sh N
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench$N rc=$?"
python - <<PY
import json
j=json.loads(open('gpurun_out/n${N}_bench.json').read().strip().split('\n')[-1])
print({k:j[k] for k in ('value','n_gpus','ms_per_step','scaling')}, 'e2e', j['e2e']['value'], j['clocks'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 50 --warmup 5 2>/dev/null | tail -c 400
