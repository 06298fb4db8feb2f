#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   — retries while the pod answers "transient / busy" (nothing charged)
T=$1; shift
for i in $(seq 1 14); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
