#!/bin/bash
set -x
TAG=${1:-r2g}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/ncu_pass.py 600 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launch_summary.txt
PPM_LANES=1 timeout 120 python tools/timeline.py 2>&1 | tee gpurun_out/${TAG}_timeline.txt
for L in 1 2; do
  echo "== $L lane(s)" | tee -a gpurun_out/${TAG}_regimes.txt
  PPM_LANES=$L timeout 200 python tools/schedule_regimes.py 2>&1 | tee -a gpurun_out/${TAG}_regimes.txt
done
