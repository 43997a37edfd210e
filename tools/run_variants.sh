#!/bin/bash
# A/B of tuning builds (tools/build_variants.sh): isolated kernel times (ncu launch list of one 1080p pass at pass 600)
# and the two-lane whole-pass time at two points of the schedule.  usage: tools/run_variants.sh name1 name2 ...  ("" = default build)
for v in "" "$@"; do
  if [ -n "$v" ]; then export PPM_B200_LIB=$PWD/ppmpa_b200/variants/libppm_b200_$v.so; else unset PPM_B200_LIB; fi
  echo "== variant '${v:-default}'"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/variant_launches_$v.csv python tools/ncu_pass.py 600 2 > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/variant_launches_$v.csv | grep -E "serialised|k_eye_expand|k_trace_photons|k_direct_light|k_dl_classify|k_gather<"
  PPM_LANES=2 timeout 200 python tools/schedule_regimes.py 1920 1080 10 2>&1 | grep -E "passes +(0|300|990)" | cut -c1-100
done
