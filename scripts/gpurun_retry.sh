#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <timeout_s> <command...>   -- retries while the pod answers "busy" (exit 3)
LOG=$1; TMO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[retry] finished rc=$rc after $i tries" >> $LOG; exit $rc; fi
  sleep 90
done
