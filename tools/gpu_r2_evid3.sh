#!/bin/bash
# round-2 final evidence on one GPU (outputs kept small: ncu reports are exported to CSV on the box and deleted)
mkdir -p gpurun_out
SECONDS=0; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$? wall ${SECONDS}s"
SECONDS=0; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_reference.json 2> gpurun_out/r2f_reference.err; echo "reference rc=$? wall ${SECONDS}s"
CMD="python bench.py --steps 64 --warmup 3 --instances 4096 --no-latency --no-sustained --no-cpu-baseline --no-parity --no-cfg4 --no-host-ceiling --no-irsplit --no-class-api --no-roofline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fwd0|k_inv0|k_mac|k_tfwd|k_tinv|k_tcols|k_trows' -s 9097 -c 704 --csv --log-file gpurun_out/r2f_launches_k4096.csv $CMD > gpurun_out/r2f_ncu_launches.log 2>&1; echo "launch list rc=$?"
export CA_TIERS=1
P="python tools/probe.py 4096 8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mac_p' -s 2280 -c 3 -o /tmp/r2f_prof_mac $P > gpurun_out/r2f_ncu_mac.log 2>&1; echo "mac rc=$?"
ncu -i /tmp/r2f_prof_mac.ncu-rep --page raw --csv > gpurun_out/r2f_mac_tiers_k4096_ncu_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'x2|k_tcols' -s 6080 -c 8 -o /tmp/r2f_prof_fft $P > gpurun_out/r2f_ncu_fft.log 2>&1; echo "fft rc=$?"
ncu -i /tmp/r2f_prof_fft.ncu-rep --page raw --csv > gpurun_out/r2f_fft_x2_k4096_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
