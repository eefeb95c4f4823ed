#!/bin/bash
# gpurun on N GPUs with retries while the pod is busy.  usage: tools/grunN.sh <N> <timeout_s> '<command>'
N=$1; T=$2; shift; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3
