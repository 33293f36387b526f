#!/bin/bash
# Round-2 evidence in one gpurun call: smoke, reference arm, ncu launch list of one tier cycle of the bench's
# timed region, ncu --set full of the three MAC launches of one period (each ncu pass after the same command
# exited 0 without it).
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 200 --warmup 10 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 64 --warmup 3 --instances 4096 --no-latency --no-sustained --no-cpu-baseline --no-roofline --no-parity --no-cfg4 --no-host-ceiling --no-irsplit --no-class-api"
timeout 300 $CMD > gpurun_out/r2e_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fwd0|k_inv0|k_mac|k_tfwd|k_tinv|k_tcols|k_trows|k_forward|k_inverse|k_tier' -s 9097 -c 704 --csv --log-file gpurun_out/r2e_launches.csv $CMD > gpurun_out/r2e_ncu1.log 2>&1
echo "launch list rc=$?"; tail -1 gpurun_out/r2e_plain.log | cut -c1-200
export CA_TIERS=1
CMD2="python tools/probe.py 4096 8"
timeout 200 $CMD2 > gpurun_out/r2e_plain_macp.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mac_p' -s 2295 -c 3 -o gpurun_out/r2e_prof_macp $CMD2 > gpurun_out/r2e_ncu2.log 2>&1
echo "mac capture rc=$?"; tail -2 gpurun_out/r2e_ncu2.log
ls -la gpurun_out | grep r2e_
