#!/bin/bash
mkdir -p gpurun_out
for b in tools/build/tc_*; do
  v=$(basename $b); echo "=== $v"; timeout 60 $b > gpurun_out/trace_$v.log 2>&1; echo "rc=$?"; grep "us/launch\|check\|lifetime" gpurun_out/trace_$v.log
done
grep -m1 -A10 "cta 0" gpurun_out/trace_tc_auto.log
