#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers busy/transient (nothing charged)
log=$1; shift; to=$1; shift
for attempt in $(seq 1 80); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|nothing was charged" $log && ! grep -q "exit code" $log; then sleep 60; continue; fi
  break
done
tail -70 $log
