#!/bin/bash
# Round-2 evidence run on one B200: tests, smoke, bench (both arms), ncu launch lists (stream-mode pass and the bench
# command itself), ncu --set full of the top kernels, timeline, regimes along the schedule, compute-sanitizer, soak.
# Everything goes to gpurun_out/r2_*; the summaries are copied to profiles/ by hand.
set -x
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r2_box.txt; nproc >> $O/r2_box.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 > $O/r2_pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke > $O/r2_smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2_bench.json 2> $O/r2_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload config2 --no-cpu > $O/r2_bench_config2.json 2>> $O/r2_bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $O/r2_reference_arm.json 2> $O/r2_reference_arm.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_pass600.csv python tools/ncu_pass.py 600 2 > /dev/null 2>&1
python tools/launch_summary.py $O/r2_launches_pass600.csv > $O/r2_launch_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-crosscheck > $O/r2_bench_under_ncu.log 2>&1
PPM_LANES=1 timeout 120 python tools/timeline.py > $O/r2_timeline_1lane.txt 2>&1
for L in 1 2; do echo "== $L lane(s)" >> $O/r2_regimes.txt; PPM_LANES=$L timeout 200 python tools/schedule_regimes.py >> $O/r2_regimes.txt 2>&1; done
echo "== config 2 (1024^2), 2 lanes" >> $O/r2_regimes.txt; timeout 200 python tools/schedule_regimes.py 1024 1024 20 >> $O/r2_regimes.txt 2>&1
timeout 900 bash tools/r2f_call.sh r2 > $O/r2_ncu_full_call.log 2>&1
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool" >> $O/r2_sanitize.txt
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py 2>&1 | tail -6 >> $O/r2_sanitize.txt
done
timeout 600 python tools/soak.py > $O/r2_soak.txt 2>&1
# BVH path: traversal throughput / pass times up to 1 M triangles, launch list and ncu --set full of a mesh-scene pass
timeout 600 python tools/bvh_bench.py --sizes 16x32,64x128,256x512,512x1024 > $O/r2_bvh_bench.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_bvh_launches.csv python tools/ncu_pass_bvh.py > /dev/null 2>&1
python tools/launch_summary.py $O/r2_bvh_launches.csv > $O/r2_bvh_launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_direct_light|k_eye_expand|k_dl_classify|k_trace_photons' --launch-skip 8 --launch-count 4 -o $O/r2_bvh_full python tools/ncu_pass_bvh.py > $O/r2_bvh_ncu.log 2>&1
ls -la $O/r2_*
