#!/bin/bash
# one gpurun call: tests -> smoke -> bench (both arms) -> variant sweep -> ncu (launch list + MAC capture)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^.\[3" | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 200 --warmup 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
for v in 0 1 2 3 4 5; do CA_MAC_VARIANT=$v timeout 300 python tools/probe.py 1024 50 2>&1 | tail -1; done
for s in 2 5; do timeout 300 python tools/probe.py 1024 50 $s 2>&1 | tail -1; done
