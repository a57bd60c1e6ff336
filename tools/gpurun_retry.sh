#!/bin/bash
# usage: [GPUS=2] tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "transient" (no slot free)
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out" | tail -60
  exit 0
done
echo "gpurun: still no slot after 40 tries"
exit 3
