#!/bin/bash
# ncu --set full of the second stream-mode pass' kernels (pass 601 of the schedule at 1080p); TAG = output prefix
TAG=${1:-r2o}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_direct_light|k_eye_expand|k_gather|k_dl_classify|k_trace_photons|k_query_mark|k_query_count|k_query_scatter|k_rs_scatter|k_combine|k_map_scatter|k_photon_place' --launch-skip 15 --launch-count 15 -o gpurun_out/${TAG}_full python tools/ncu_pass.py 600 2 > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
