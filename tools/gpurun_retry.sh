#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'   (retries while the pod answers busy/transient)
T=$1; shift
mkdir -p gpurun_out
for attempt in $(seq 1 80); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|retry in a few minutes\|another call"; then
    echo "[retry $attempt $(date +%H:%M:%S)] busy/transient (rc=$rc)" >> gpurun_out/retry.log; sleep 45; continue
  fi
  echo "$out"; exit $rc
done
echo "gave up after 80 attempts"; exit 3
