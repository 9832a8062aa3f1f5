#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers busy/transient (nothing charged)
log=$1; shift; to=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|nothing was charged" $log && ! grep -q "exit code" $log; then sleep 120; continue; fi
  break
done
tail -70 $log
