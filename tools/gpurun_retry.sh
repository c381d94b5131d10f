#!/bin/bash
# local helper: gpurun with retries while the pod answers "busy" (exit code 3, nothing charged)
# usage: tools/gpurun_retry.sh <timeout_s> <logfile> '<command>' [--gpus N]
T=$1; LOG=$2; CMD=$3; shift 3
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
