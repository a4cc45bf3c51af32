#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -q -m gpu --timeout=600 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/gpu_tests.log
for v in "ffma ldg" "tc ldg"; do
  set -- $v
  SPK_B200_DENSE=$1 SPK_B200_EDGE=$2 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
n = f"gpurun_out/bench_{sys.argv[1]}_{sys.argv[2]}"
try:
    d = json.load(open(n + ".json"))
    print(sys.argv[1:], "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:(round(v["avg_us"],1), round(v["frac"],3)) for k,v in d["roofline_all"].items()}, "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "eager", d["eager_gpu_baseline"])
except Exception as e:
    print(sys.argv[1:], "failed", e); print(open(n + ".err").read()[-800:])
PY
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
