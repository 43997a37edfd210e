#!/bin/bash
# ncu: launch list of two stream-mode passes, and --set full of the second pass' top kernels
set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_direct_light|k_eye_expand|k_gather|k_dl_classify|k_trace_photons|k_query_mark|k_combine' --launch-skip 10 --launch-count 10 -o gpurun_out/r2f_full python tools/ncu_pass.py 600 2 > gpurun_out/r2f_ncu2.log 2>&1
ls -la gpurun_out/r2f*
