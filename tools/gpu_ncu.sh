#!/bin/bash
# ncu evidence: launch list of the timed step + one full capture of the dominant kernel (k_mac).
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --instances 1024 --no-latency --no-sustained --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_forward|k_mac|k_inverse' -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_mac -s 4 -c 1 -o gpurun_out/prof_mac $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out/
