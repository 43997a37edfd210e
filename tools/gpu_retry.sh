#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <command string>   -- retries while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > gpurun_out/.retry_last.txt 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" gpurun_out/.retry_last.txt || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
cat gpurun_out/.retry_last.txt
