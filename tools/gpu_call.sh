#!/bin/bash
# usage: tools/gpu_call.sh <log name> <timeout s> [gpurun extra args] -- retries while the pod answers busy (exit 3)
name=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" -- 'bash tools/_run.sh' > gpurun_out/$name.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/$name.log; then break; fi
  sleep 60
done
echo "done rc=$rc" >> gpurun_out/$name.log
