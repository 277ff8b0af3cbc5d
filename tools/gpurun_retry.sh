#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3): gpurun_retry.sh TIMEOUT 'command'
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
