#!/bin/bash
# round-2 (16 x 16 rows, growth-4 batches): GPU suite, ncu --set full of the MAC launches and of the FFT kernels of one period (K = 4096)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
export CA_TIERS=1
CMD="python tools/probe.py 4096 8"
timeout 200 $CMD > gpurun_out/plain_x2.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/plain_x2.log | cut -c1-330
# 760 periods before the timed loop x 4 MAC launches
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mac_p' -s 3040 -c 4 -o gpurun_out/prof_macp_g4 $CMD > gpurun_out/ncu_macp_g4.log 2>&1; echo "mac rc=$?"
# FFT kernels per period: fwd0, inv0, tier1 fwd/inv, tier2 fwd/inv, tier3 cols+rows fwd/inv = 10
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'x2|k_tcols' -s 7600 -c 10 -o gpurun_out/prof_fft_x2 $CMD > gpurun_out/ncu_fft_x2.log 2>&1; echo "fft rc=$?"
ls -la gpurun_out/*.ncu-rep
