#!/bin/bash
# 2-GPU validation: NCCL reduce inside the C ABI (python ranks, C++ ranks), bench at N=2
set -x
nvidia-smi -L | tee gpurun_out/r2j_box.txt
timeout 600 python -m pytest tests -m gpu -q -k "two_gpu or ppmpa_frame or sharded" --timeout 300 2>&1 | tail -15 | tee gpurun_out/r2j_pytest_2gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
tail -3 gpurun_out/r2j_bench_n2.err
PPM_SEED=1 timeout 300 ./ppmpa_b200/bin/ppmpa_frame -g 2 100 1000000 0.1 examples/camera0.scr examples/ex-glassbox.scene gpurun_out/r2j_frame.exr 2>&1 | tee gpurun_out/r2j_frame.txt
