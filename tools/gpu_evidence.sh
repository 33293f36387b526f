#!/bin/bash
# Round evidence in one gpurun call: tests, smoke, both bench arms, ncu launch list + full MAC capture.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed" > gpurun_out/ev_tests.txt; cat gpurun_out/ev_tests.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 200 --warmup 10 > gpurun_out/ev_bench_ref.json 2> gpurun_out/ev_bench_ref.err; echo "ref rc=$?"
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/ev_bench_ours.json 2> gpurun_out/ev_bench_ours.err; echo "ours rc=$?"
CMD="python bench.py --steps 8 --warmup 3 --instances 4096 --no-latency --no-sustained --no-cpu-baseline --no-roofline"
timeout 200 $CMD > gpurun_out/ev_plain.log 2>&1 &&
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_forward|k_mac|k_inverse|k_tier_forward|k_tier_inverse' -s 7560 -c 144 --csv --log-file gpurun_out/ev_launches.csv $CMD > gpurun_out/ev_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 200 $CMD > gpurun_out/ev_plain2.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_mac -s 2520 -c 3 -o gpurun_out/ev_prof_mac_tiers $CMD > gpurun_out/ev_ncu2.log 2>&1
echo "mac capture rc=$?"
ls -la gpurun_out | grep ev_
