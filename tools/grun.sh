#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 = nothing charged).  usage: tools/grun.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 45
done
exit 3
