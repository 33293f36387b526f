#!/bin/bash
# ncu --set full of every FFT kernel of one tiered period at K = 4096 (after the plain run exits 0)
mkdir -p gpurun_out
export CA_TIERS=1
CMD="python tools/probe.py 4096 8"
timeout 120 $CMD > gpurun_out/plain_fft.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_tier_forward|k_tier_inverse|k_forward|k_inverse' -s 4580 -c 6 -o gpurun_out/prof_fft $CMD > gpurun_out/ncu_fft.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_fft.log
