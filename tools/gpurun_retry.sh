#!/bin/bash
# gpurun with retries while the pod answers "busy" (status=transient / exit 3): tools/gpurun_retry.sh <timeout-seconds> <log> <command...>
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > $LOG 2>&1
  if grep -q "status=transient" $LOG; then sleep 90; continue; fi
  break
done
