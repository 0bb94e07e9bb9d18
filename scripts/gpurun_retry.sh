#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers "busy" (exit 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
