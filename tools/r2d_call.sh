#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -s --timeout 300 2>&1 | tail -40 | tee gpurun_out/r2d_pytest_configs.txt
for L in 1 2; do
  echo "== graph mode + node priorities, $L lane(s)" | tee -a gpurun_out/r2d_regimes.txt
  PPM_LANES=$L timeout 200 python tools/schedule_regimes.py 2>&1 | tee -a gpurun_out/r2d_regimes.txt
done
