#!/bin/bash
# One-shot evidence run on a B200 box (gpurun): bench line, ncu launch list, ncu --set full of the three
# FP64-heavy kernels, GPU parity suite, reference arm, smoke.  Every step is bounded and writes into gpurun_out/
# as it goes, most valuable first, so a clamped call still brings the early files back.
#   gpurun --timeout 780 -- 'bash tools/r1d_capture.sh r1d'
TAG=${1:-r1d}
OUT=gpurun_out
mkdir -p $OUT
cd "$(dirname "$0")/.."
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $OUT/${TAG}_log.txt; }
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_box.txt 2>&1
nproc >> $OUT/${TAG}_box.txt

stamp "bench (N=1, default steps)"
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
stamp "bench rc=$? $(head -c 300 $OUT/${TAG}_bench.json)"

stamp "ncu launch list"
PPM_LANES=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu > $OUT/${TAG}_launches_bench.txt 2>&1
stamp "launch list rc=$?"

stamp "ncu --set full: k_gather, k_direct_light, k_dl_classify (pass 14 of the schedule)"
PPM_LANES=1 timeout 240 ncu --set full --clock-control none --import-source on \
  -k 'regex:^k_gather$|^k_direct_light$|^k_dl_classify$' -s 42 -c 3 -f -o $OUT/${TAG}_gather_dl \
  python tools/pass_phases.py 16 > $OUT/${TAG}_ncu_full.txt 2>&1
stamp "ncu full rc=$?"
ncu -i $OUT/${TAG}_gather_dl.ncu-rep --page raw --csv > $OUT/${TAG}_gather_dl_ncu_raw.csv 2>> $OUT/${TAG}_ncu_full.txt

stamp "pytest -m gpu"
timeout 420 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.txt 2>&1
stamp "pytest rc=$? $(tail -n 1 $OUT/${TAG}_pytest_gpu.txt)"

stamp "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1
stamp "smoke rc=$? $(tail -n 1 $OUT/${TAG}_smoke.txt)"

stamp "reference arm"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_reference_arm.json 2> $OUT/${TAG}_reference_arm.err
stamp "reference arm rc=$?"

stamp "phase times, single lane"
PPM_LANES=1 timeout 100 python tools/pass_phases.py 14 > $OUT/${TAG}_pass_phases.txt 2>&1
stamp "done"
