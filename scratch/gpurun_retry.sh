#!/bin/bash
# usage: scratch/gpurun_retry.sh <timeout_s> [--gpus N] -- <command>   : retries while the pod answers "no slot" (exit 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
