#!/bin/bash
TAG=${1:-r2w}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/ncu_pass.py 600 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv 2>/dev/null | head -8 | tee gpurun_out/${TAG}_launch_summary.txt
PPM_DL_STATS=1 PPM_LANES=1 timeout 60 python tools/pass_phases.py 2 2>&1 | grep -m 2 "ppm direct light" | tee gpurun_out/${TAG}_dl_stats.txt
bash tools/bench_variants.sh 2>&1 | tee gpurun_out/${TAG}_bench.txt
