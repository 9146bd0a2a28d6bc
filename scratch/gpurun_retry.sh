#!/bin/bash
# usage: gpurun_retry.sh <outfile> <gpurun args...>   — retries while the pod is busy (exit 3)
out=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
