#!/bin/bash
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2e_box.txt
PPM_LANES=1 timeout 120 python tools/timeline.py 2>&1 | tee gpurun_out/r2e_timeline.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 | tee gpurun_out/r2e_pytest_gpu.txt
