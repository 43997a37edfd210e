#!/bin/bash
N=$(nvidia-smi -L | wc -l)
O=gpurun_out
lscpu | grep -i "numa\|socket\|model name" > $O/r2_n${N}_numa.txt; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo "$d $(cat $d/numa_node) $(cat $d/local_cpulist)"; fi; done >> $O/r2_n${N}_numa.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > $O/r2b_bench_n${N}.json 2> $O/r2b_bench_n${N}.err
python -c "
import json; d=json.load(open('$O/r2b_bench_n${N}.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
cat $O/r2_n${N}_numa.txt
