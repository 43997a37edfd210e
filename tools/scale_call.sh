#!/bin/bash
# bench at N = visible GPUs through torchrun (the driver's launch line)
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n${N}.json').read().strip().splitlines()[-1]); print('N=%d: %.3f ms/step, %.3f G pixels/s, e2e %.3f G, gather frac %.3f' % (d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['frac']))"
