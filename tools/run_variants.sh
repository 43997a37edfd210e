#!/bin/bash
# On the GPU box: bench every ppmpa_b200/variants/libppm_b200_*.so (and the default build) with the same short run.
#   gpurun -- 'bash tools/run_variants.sh [steps]'
cd "$(dirname "$0")/.."
STEPS=${1:-60}
mkdir -p gpurun_out
out=gpurun_out/variants.txt
: > $out
run() {  # name, lib
  for rep in $(seq 1 ${REPS:-2}); do
    line=$(PPM_B200_LIB=$2 timeout 120 python bench.py --steps $STEPS --warmup 3 --no-cpu 2>/dev/null)
    echo "$1 rep$rep $(echo "$line" | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_pass']; print('ms_per_step=%.4f e2e=%.1fM gatherk=%.3f dl=%.3f expand=%.3f trace=%.3f build=%.3f frac=%.3f' % (d['ms_per_step'], d['e2e']['value']/1e6, p['gather_kernel'], p['direct_light'], p['eye_expand'], p['photon_trace'], p['map_build'], d['roofline']['frac']))" 2>&1 | tail -n 1)" | tee -a $out
  done
}
run default ""
for lib in ppmpa_b200/variants/libppm_b200_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#libppm_b200_}
  run $n $PWD/$lib
done
run default_again ""
