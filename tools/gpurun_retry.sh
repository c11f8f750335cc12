#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> <timeout_s> '<command>'   -> gpurun_out/call_<tag>.txt ; retries while the pod answers busy (rc 3)
tag=$1; to=$2; shift 2
mkdir -p gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$to" -- "$@" > gpurun_out/call_$tag.txt 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
