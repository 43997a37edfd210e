#!/bin/bash
# Round-2 first GPU call: validate and measure the candidates that were written without GPU access at the end of round 1
# (DESIGN.md section 11).  Build them HERE first (no GPU needed):
#   tools/build_variants.sh dlregion:-DPPM_DL_REGION=1 pipe:-DGATHER_PIPE=1 "both:-DPPM_DL_REGION=1 -DGATHER_PIPE=1"
# then:  gpurun --timeout 600 -- 'bash tools/run_candidates.sh'
# Every variant must pass the WHOLE GPU parity suite (bit-exact culling test incl. the cell-sorted probe order, oracle
# comparisons, golden fixtures) before its numbers mean anything.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in ppmpa_b200/variants/libppm_b200_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#libppm_b200_}
  echo "== pytest -m gpu with $n"
  PPM_B200_LIB=$PWD/$lib timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 | tee gpurun_out/cand_${n}_pytest.txt
done
REPS=2 bash tools/run_variants.sh 100
cp gpurun_out/variants.txt gpurun_out/cand_variants.txt
echo "== radius-schedule regimes (configs[4]), default build"
timeout 120 python tools/schedule_regimes.py | tee gpurun_out/cand_regimes_default.txt
for n in pipe both; do
  lib=ppmpa_b200/variants/libppm_b200_$n.so
  [ -e "$lib" ] || continue
  echo "== radius-schedule regimes (configs[4]), $n"
  PPM_B200_LIB=$PWD/$lib timeout 120 python tools/schedule_regimes.py | tee gpurun_out/cand_regimes_$n.txt
done
echo "== culling statistics of the region classifier (tested primitives per node), config 2"
for n in dlregion; do
  lib=ppmpa_b200/variants/libppm_b200_$n.so
  [ -e "$lib" ] || continue
  PPM_DL_STATS=1 PPM_LANES=1 PPM_B200_LIB=$PWD/$lib timeout 60 python tools/pass_phases.py 3 2>&1 | grep -m 4 "ppm direct light" | tee gpurun_out/cand_dlregion_stats.txt
done
PPM_DL_STATS=1 PPM_LANES=1 timeout 60 python tools/pass_phases.py 3 2>&1 | grep -m 4 "ppm direct light" | tee gpurun_out/cand_default_stats.txt
