#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> '<command>' [extra gpurun args]; retries while the pod answers "transient" / busy (rc 3)
T=$1; shift; CMD=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|nothing was charged"; then sleep 90; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up: pod busy"; exit 3
