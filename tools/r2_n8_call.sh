#!/bin/bash
# N-GPU evidence (N = number of visible GPUs): the multi-GPU tests, bench through torchrun, and the whole north-star
# job from the C++ host
set -x
N=$(nvidia-smi -L | wc -l)
O=gpurun_out
nvidia-smi -L > $O/r2_n${N}_box.txt; nproc >> $O/r2_n${N}_box.txt
timeout 600 python -m pytest tests -m gpu -q -k "two_gpu or ppmpa_frame or sharded" --timeout 300 2>&1 | tail -5 | tee $O/r2_pytest_${N}gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > $O/r2_bench_n${N}.json 2> $O/r2_bench_n${N}.err
tail -2 $O/r2_bench_n${N}.err
PPM_SEED=1 timeout 600 ./ppmpa_b200/bin/ppmpa_frame -g $N 1000 1000000 0.1 examples/camera_1080p.scr examples/ex-glassbox.scene $O/r2_frame_n${N}.ppm 2>&1 | grep -v "^NCCL" | tee $O/r2_ppmpa_frame_n${N}.txt
head -c 300 $O/r2_frame_n${N}.ppm | head -5; rm -f $O/r2_frame_n${N}.ppm
