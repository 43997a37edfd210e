#!/bin/bash
# Builds tuning variants of libppm_b200.so (compile-time macros) into ppmpa_b200/variants/ (git-ignored, travels to the
# GPU box).  tools/run_variants.sh benches each one there through PPM_B200_LIB.
#   tools/build_variants.sh name1:-DFOO=1 name2:"-DBAR=2 -DBAZ=3" ...
set -e
cd "$(dirname "$0")/../ppmpa_b200/csrc"
make -s all
mkdir -p ../variants build
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  (
    $NVCC -O3 -std=c++17 -lineinfo -fmad=false $ARCH -Xcompiler -fPIC -Xcompiler -ffp-contract=off -Xcompiler -pthread $defs \
      -Xptxas -v -c engine.cu -o build/engine_$name.o 2> build/engine_$name.ptxas.log
    $NVCC $ARCH -shared -cudart static -Xcompiler -pthread -o ../variants/libppm_b200_$name.so build/engine_$name.o build/host_model.o build/host_parse.o build/host_io.o build/host_bvh.o -ldl
    echo "built $name ($defs)"
  ) &
  while [ $(jobs -r | wc -l) -ge 6 ]; do sleep 0.5; done
done
wait
