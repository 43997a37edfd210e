#!/bin/bash
# Round 2, GPU call A: validate the two candidates written blind at the end of round 1 (GATHER_PIPE, PPM_DL_REGION)
# and take single-lane phase numbers along the configs[4] schedule as the baseline for this round's work.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/r2a_box.txt
for n in dlregion pipe; do
  lib=ppmpa_b200/variants/libppm_b200_$n.so
  [ -e "$lib" ] || continue
  echo "== pytest -m gpu with $n"
  PPM_B200_LIB=$PWD/$lib timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 | tee gpurun_out/r2a_${n}_pytest.txt
done
echo "== schedule regimes, two lanes (default)"
timeout 200 python tools/schedule_regimes.py 2>&1 | tee gpurun_out/r2a_regimes_default.txt
for n in pipe dlregion both; do
  echo "== schedule regimes, two lanes ($n)"
  PPM_B200_LIB=$PWD/ppmpa_b200/variants/libppm_b200_$n.so timeout 200 python tools/schedule_regimes.py 2>&1 | tee gpurun_out/r2a_regimes_$n.txt
done
echo "== schedule regimes, ONE lane (serialised phases), default and both"
PPM_LANES=1 timeout 200 python tools/schedule_regimes.py 2>&1 | tee gpurun_out/r2a_regimes_default_1lane.txt
PPM_LANES=1 PPM_B200_LIB=$PWD/ppmpa_b200/variants/libppm_b200_both.so timeout 200 python tools/schedule_regimes.py 2>&1 | tee gpurun_out/r2a_regimes_both_1lane.txt
echo "== direct-light culling statistics"
PPM_DL_STATS=1 PPM_LANES=1 PPM_B200_LIB=$PWD/ppmpa_b200/variants/libppm_b200_dlregion.so timeout 60 python tools/pass_phases.py 3 2>&1 | grep -m 4 "ppm direct light" | tee gpurun_out/r2a_dlregion_stats.txt
PPM_DL_STATS=1 PPM_LANES=1 timeout 60 python tools/pass_phases.py 3 2>&1 | grep -m 4 "ppm direct light" | tee gpurun_out/r2a_default_stats.txt
