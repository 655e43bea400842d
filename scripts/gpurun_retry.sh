#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
#   scripts/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]
t=$1; cmd=$2; gpus=${3:-1}
for i in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$t" -- "$cmd"; else /usr/local/graft/bin/gpurun --gpus "$gpus" --timeout "$t" -- "$cmd"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
