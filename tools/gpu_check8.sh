#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -q -m gpu --timeout=300 -x > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/gpu_tests.log
for e in async ldg; do
SPK_B200_EDGE=$e timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$e.json 2> gpurun_out/bench_$e.err
python - $e <<'PY'
import json, sys
n = "gpurun_out/bench_" + sys.argv[1]
try:
    d = json.load(open(n + ".json"))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:(round(v["avg_us"],1), round(v["frac"],3), round(v["share_of_step"],3)) for k,v in d["roofline_all"].items()}, "launches", d.get("gpu_launches"))
except Exception as e:
    print("failed", e); print(open(n + ".err").read()[-1500:])
PY
done
