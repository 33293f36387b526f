#!/bin/bash
mkdir -p gpurun_out
export CA_TIERS=1
CMD="python tools/probe.py 2048 64"
$CMD > gpurun_out/plain_t.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_forward|k_mac|k_inverse|k_tier_forward|k_tier_inverse' -s 6930 -c 576 --csv --log-file gpurun_out/launches_tiers.csv $CMD > gpurun_out/ncu_t.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain_t.log | cut -c1-200
