#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'   (retries while the pod answers busy/transient)
T=$1; shift
for attempt in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|retry in a few minutes"; then
    echo "[retry $attempt] busy/transient (rc=$rc); sleeping 120 s"; sleep 120; continue
  fi
  echo "$out"; exit $rc
done
echo "gave up after 20 attempts"; exit 3
