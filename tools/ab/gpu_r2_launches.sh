#!/bin/bash
# ncu launch list of one tier cycle (64 periods) of the bench's device-timed region at K = 4096, plus host-pipeline chunk sweep
mkdir -p gpurun_out
CMD="python bench.py --steps 64 --warmup 3 --instances 4096 --no-latency --no-sustained --no-cpu-baseline --no-roofline --no-parity --no-cfg4 --no-host-ceiling --no-irsplit --no-class-api"
timeout 300 $CMD > gpurun_out/r2e_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fwd0|k_inv0|k_mac|k_tfwd|k_tinv|k_tcols|k_trows' -s 9097 -c 704 --csv --log-file gpurun_out/r2e_launches.csv $CMD > gpurun_out/r2e_ncu1.log 2>&1
echo "launch list rc=$?"
export CA_TIERS=1 CA_NOPROFILE=1
for c in 1 2 3 4; do echo "== chunks=$c K=16128"; CA_IO_CHUNKS=$c timeout 600 python tools/probe.py 16128 128 2>&1 | tail -1 | cut -c1-200; done
