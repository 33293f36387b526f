#!/bin/bash
# dynamic work hand-out in the persistent MAC: first under bounded mbarrier spins (libD = -DCA_MBAR_DEBUG), then release
mkdir -p gpurun_out
CA_B200_LIB=$PWD/gpurun_tmp/libD.so timeout 300 python -m pytest tests/test_tiers_gpu.py -m gpu -x -q -k "persistent or pipelined or staggered or schedule_bits" > gpurun_out/r2f_dbg.log 2>&1; echo "debug-lib tests rc=$?"; tail -5 gpurun_out/r2f_dbg.log | cut -c1-300; grep -c "mbar timeout" gpurun_out/r2f_dbg.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2f_pytest.log | cut -c1-300
export CA_TIERS=1
echo "== profile K=4096"; timeout 300 python tools/probe.py 4096 64 2>&1 | tail -2 | cut -c1-400
echo "== noprofile K=4096"; CA_NOPROFILE=1 timeout 300 python tools/probe.py 4096 128 2>&1 | tail -1 | cut -c1-200
echo "== noprofile K=16128"; CA_NOPROFILE=1 timeout 600 python tools/probe.py 16128 128 2>&1 | tail -1 | cut -c1-200
