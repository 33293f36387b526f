#!/bin/bash
# ncu --set full of every row-FFT kernel of one tiered period at K = 4096 (after the plain run exits 0)
mkdir -p gpurun_out
export CA_TIERS=1
CMD="python tools/probe.py 4096 8"
timeout 120 $CMD > gpurun_out/plain_rows.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rows|k_tfwd|k_tinv|k_tcols|k_trows' -s 6096 -c 8 -o gpurun_out/prof_rows $CMD > gpurun_out/ncu_rows.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_rows.log; tail -2 gpurun_out/plain_rows.log | cut -c1-300
