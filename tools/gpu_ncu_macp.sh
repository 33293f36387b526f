#!/bin/bash
# ncu --set full of the persistent MAC launches of one tiered period at K = 4096
mkdir -p gpurun_out
export CA_TIERS=1
CMD="python tools/probe.py 4096 8"
timeout 120 $CMD > gpurun_out/plain_macp.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_mac_p' -s 2295 -c 3 -o gpurun_out/prof_macp $CMD > gpurun_out/ncu_macp.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_macp.log
