#!/bin/bash
# usage: tools/gpu.sh <log> <timeout-seconds> <command...>   -- gpurun with retries while the pod answers "transient"
LOG=$1; shift; TMO=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then break; fi
  sleep 90
done
tail -5 $LOG
