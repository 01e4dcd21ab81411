#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <log> [--gpus N] -- '<command>'   — retries while gpurun answers "busy" (rc 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> $LOG; exit $rc; fi
  sleep 90
done
echo "gave up" >> $LOG; exit 3
