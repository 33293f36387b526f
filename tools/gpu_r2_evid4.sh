#!/bin/bash
# ncu launch list of one tier cycle (64 periods) of the bench's device-timed region at K = 4096
mkdir -p gpurun_out
CMD="python bench.py --steps 64 --warmup 3 --instances 4096 --no-clocks --no-latency --no-sustained --no-cpu-baseline --no-parity --no-cfg4 --no-host-ceiling --no-irsplit --no-class-api --no-roofline"
timeout 120 $CMD > gpurun_out/r2f_plain_k4096.json 2> gpurun_out/r2f_plain_k4096.err; echo "plain rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fwd0|k_inv0|k_mac|k_tfwd|k_tinv|k_tcols|k_trows' -s 9097 -c 704 --csv --log-file gpurun_out/r2f_launches_k4096.csv $CMD > gpurun_out/r2f_ncu_launches.log 2>&1; rc=$?; echo "launch list rc=$rc"
if [ $rc -ne 0 ]; then
  export CA_TIERS=1 CA_NOPROFILE=1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fwd0|k_inv0|k_mac|k_tfwd|k_tinv|k_tcols|k_trows' -s 8360 -c 704 --csv --log-file gpurun_out/r2f_launches_probe_k4096.csv python tools/probe.py 4096 64 > gpurun_out/r2f_ncu_launches2.log 2>&1; echo "probe launch list rc=$?"
fi
wc -l gpurun_out/r2f_launches*.csv
