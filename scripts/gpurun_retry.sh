#!/bin/bash
# bash scripts/gpurun_retry.sh <log> <timeout-seconds> <command...>: retries while the pod answers "busy" (exit 3; nothing is charged).
LOG=$1; TMO=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $attempt)" >> $LOG; exit $rc; fi
  sleep 120
done
echo "gpurun: still busy after 40 attempts" >> $LOG; exit 3
