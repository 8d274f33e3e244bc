#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout> <command...>   — retries while the pod answers "busy" (exit 3)
log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
exit 3
