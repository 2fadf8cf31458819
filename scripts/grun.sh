#!/bin/bash
# gpurun with retry on "no box free" (exit 3).  usage: [GPUS=N] scripts/grun.sh <timeout_s> '<command>'
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$t" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
