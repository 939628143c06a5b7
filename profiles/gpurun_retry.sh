#!/bin/bash
# gpurun with retries while the pod has no box free (exit code 3: nothing charged).  Usage: gpurun_retry.sh <log> <timeout> '<command>' [gpus]
log=$1; to=$2; cmd=$3; gpus=${4:-1}
for i in $(seq 1 20); do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$cmd" > $log 2>&1; else /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$cmd" > $log 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
