#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_kernels.py -q -m gpu --timeout=120 -x -k "tensor_core_filter" > gpurun_out/tcedge_test.log 2>&1; echo "tc edge test rc=$?"; tail -8 gpurun_out/tcedge_test.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
python - <<'PY'
import json
n = "gpurun_out/bench_tc"
try:
    d = json.load(open(n + ".json"))
    print("ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:(round(v["avg_us"],1), round(v["frac"],3), round(v["share_of_step"],3)) for k,v in d["roofline_all"].items()})
except Exception as e:
    print("failed", e); print(open(n + ".err").read()[-1500:])
PY
