#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <timeout> [--gpus N] -- '<command>'   : gpurun with retries while the pod answers busy (exit code 3)
log=$1; shift; to=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
  gpurun --timeout $to "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
exit 3
