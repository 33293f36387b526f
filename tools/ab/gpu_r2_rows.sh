#!/bin/bash
# row-FFT family: A/B tests first, then the whole suite, then per-kernel times new vs legacy
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiers_gpu.py -m gpu -q -k "row_fft" > gpurun_out/r2b_rows_tests.log 2>&1; echo "rows tests rc=$?"
tail -40 gpurun_out/r2b_rows_tests.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "suite rc=$?"; tail -15 gpurun_out/r2b_pytest.log | cut -c1-300
export CA_TIERS=1
for leg in "" 1; do
  echo "== legacy=$leg profile K=4096"; CA_LEGACY_FFT=$leg timeout 300 python tools/probe.py 4096 64 2>&1 | tail -2 | cut -c1-400
  echo "== legacy=$leg noprofile K=4096"; CA_LEGACY_FFT=$leg CA_NOPROFILE=1 timeout 300 python tools/probe.py 4096 128 2>&1 | tail -1 | cut -c1-400
done
