#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> <command...> — retries while the pod answers "busy" (exit 3, nothing charged)
log=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
