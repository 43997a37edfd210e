#!/bin/bash
# bench.py (north-star job, 20 strided passes) per tuning build: ms/step, gather ms/launch and logical fraction
for v in "" "$@"; do
  if [ -n "$v" ]; then export PPM_B200_LIB=$PWD/ppmpa_b200/variants/libppm_b200_$v.so; else unset PPM_B200_LIB; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-crosscheck 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('variant %-10s ms/step %.3f  gather %.4f ms/launch  logical frac %.4f  fp64 %.3f  e2e %.1f M' % ('${v:-default}', d['ms_per_step'], r['ms_per_launch'], r['frac'], r['fp64_issue_frac'], d['e2e']['value']/1e6))"
done
