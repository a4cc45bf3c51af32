#!/bin/bash
mkdir -p gpurun_out
for c in cfg4 cfg2; do
timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --md > gpurun_out/bench_md_$c.json 2> gpurun_out/bench_md_$c.err
python - $c <<'PY'
import json, sys
c = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_md_{c}.json"))
    print(c, "ms/step", round(d["ms_per_step"], 3), "md", d.get("md"))
except Exception as e:
    print(c, "failed", e); print(open(f"gpurun_out/bench_md_{c}.err").read()[-1200:])
PY
done
