#!/bin/bash
# A/B of k_eye_expand variants: ncu launch list per variant (isolated kernel time) + 2-lane whole-pass time
for v in "" oldstack eye5 old5 eye4; do
  if [ -n "$v" ]; then export PPM_B200_LIB=$PWD/ppmpa_b200/variants/libppm_b200_$v.so; else unset PPM_B200_LIB; fi
  echo "== variant '${v:-default}'"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_$v.csv python tools/ncu_pass.py 600 2 > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r2i_launches_$v.csv | grep -E "serialised|k_eye_expand|k_trace_photons"
  PPM_LANES=2 timeout 200 python tools/schedule_regimes.py 1920 1080 10 2>&1 | grep -E "passes  (300|990)" | cut -c1-60
done
