#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -q -m gpu --timeout=300 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/gpu_tests.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in cfg4 cfg2; do
timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --md > gpurun_out/bench_md_$c.json 2> gpurun_out/bench_md_$c.err
python - $c <<'PY'
import json, sys
c = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_md_{c}.json"))
    print(c, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: (round(v["avg_us"], 1), round(v["frac"], 3)) for k, v in (d["roofline_all"] or {}).items()}, "traffic", (d["roofline"] or {}).get("traffic"), "md", d.get("md"))
except Exception as e:
    print(c, "failed", e); print(open(f"gpurun_out/bench_md_{c}.err").read()[-1200:])
PY
done
