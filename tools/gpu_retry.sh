#!/bin/bash
# usage: tools/gpu_retry.sh <log> <timeout> [--gpus N] -- '<command>'   (retries while gpurun answers "busy": exit code 3)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO "$@" > $LOG 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then echo "rc=$rc" >> $LOG; exit $rc; fi
  sleep 60
done
