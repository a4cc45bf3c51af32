#!/bin/bash
mkdir -p gpurun_out
for v in base mb5 mb6; do
lib=""; [ $v != base ] && lib=$PWD/tools/build/libspk_$v.so
SPK_B200_LIB=$lib SPK_B200_EDGE=ldg timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
python - $v <<'PY'
import json, sys
n = "gpurun_out/bench_" + sys.argv[1]
try:
    d = json.load(open(n + ".json"))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:(round(v["avg_us"],1), round(v["frac"],3), round(v["share_of_step"],3)) for k,v in d["roofline_all"].items()})
except Exception as e:
    print("failed", e); print(open(n + ".err").read()[-1500:])
PY
done
