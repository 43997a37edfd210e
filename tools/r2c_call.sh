#!/bin/bash
# round 2: regimes along the schedule with the device-resident pass (graph mode), 1/2/3 lanes; stream mode for comparison
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/r2c_box.txt
timeout 300 python -m pytest tests -m gpu -q -k "cull or heavy or cli or lanes" --timeout 200 2>&1 | tail -5 | tee gpurun_out/r2c_pytest_subset.txt
for L in 1 2 3; do
  echo "== graph mode, $L lane(s)" | tee -a gpurun_out/r2c_regimes.txt
  PPM_LANES=$L timeout 200 python tools/schedule_regimes.py 2>&1 | tee -a gpurun_out/r2c_regimes.txt
done
echo "== stream mode (PPM_GRAPH=0), 2 lanes" | tee -a gpurun_out/r2c_regimes.txt
PPM_GRAPH=0 PPM_LANES=2 timeout 200 python tools/schedule_regimes.py 2>&1 | tee -a gpurun_out/r2c_regimes.txt
echo "== config 2 (1024^2), 2 lanes" | tee -a gpurun_out/r2c_regimes.txt
PPM_LANES=2 timeout 200 python tools/schedule_regimes.py 1024 1024 20 2>&1 | tee -a gpurun_out/r2c_regimes.txt
