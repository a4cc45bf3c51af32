#!/bin/bash
# edge-kernel experiments (round 2): phase trace of k_painn_edge_fwd_tc, and A/B timing of variant builds via bench.py
mkdir -p gpurun_out
C=schnetpack_b200/csrc
SPK_B200_LIB=$C/libspk_b200_trace.so timeout 300 python tools/edge_trace.py > gpurun_out/edge_trace.txt 2>&1
tail -45 gpurun_out/edge_trace.txt
for v in "" $@; do
  lib=$C/libspk_b200${v:+_$v}.so
  SPK_B200_LIB=$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-spatial > gpurun_out/edge_ab_${v:-base}.json 2> gpurun_out/edge_ab_${v:-base}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/edge_ab_${v:-base}.json"))
    print("${v:-base}", "ms/step", round(d["ms_per_step"], 4), {k: round(x["avg_us"], 1) for k, x in d["roofline_all"].items()})
except Exception as e:
    print("${v:-base} failed", e); print(open("gpurun_out/edge_ab_${v:-base}.err").read()[-800:])
PY
done
