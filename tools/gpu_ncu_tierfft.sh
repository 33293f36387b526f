#!/bin/bash
mkdir -p gpurun_out
export CA_TIERS=1
CMD="python tools/probe.py 4096 8"
timeout 120 $CMD > gpurun_out/plain_tf.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tier_forward|k_tier_inverse' -s 3050 -c 4 -o gpurun_out/prof_tierfft $CMD > gpurun_out/ncu_tf.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_tf.log
