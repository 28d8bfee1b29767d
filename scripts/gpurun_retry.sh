#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <tag> <command...>   — retries while gpurun answers "busy" (exit 3)
T=$1; TAG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > gpurun_out/run_$TAG.out 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)"; exit $rc; fi
  sleep 150
done
echo "gave up"; exit 3
