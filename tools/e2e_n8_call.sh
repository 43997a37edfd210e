#!/bin/bash
# e2e at N = visible GPUs for different numbers of host threads / contexts per GPU
N=$(nvidia-smi -L | wc -l)
for L in 4 3 2; do
  PPM_E2E_LANES=$L timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$L bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-crosscheck 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('e2e lanes $L: device %.3f ms/step %.3f G pixels/s; e2e %.3f G pixels/s (%.2f of device)' % (d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['value']/d['value']))"
done
