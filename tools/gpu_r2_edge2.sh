#!/bin/bash
# edge kernels with row-switch prefetch: numerics + timing + trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_kernels.py tests/test_cuda_parity.py -q -m gpu -x --timeout=300 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-spatial > gpurun_out/edge2_bench.json 2> gpurun_out/edge2_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/edge2_bench.json"))
print("ms/step", round(d["ms_per_step"], 4), {k: round(x["avg_us"], 1) for k, x in d["roofline_all"].items()}, d["clocks"])
PY
timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-spatial > gpurun_out/edge2_bench_cfg4.json 2> gpurun_out/edge2_bench_cfg4.err
python - <<PY
import json
d = json.load(open("gpurun_out/edge2_bench_cfg4.json"))
print("cfg4 ms/step", round(d["ms_per_step"], 4), {k: round(x["avg_us"], 1) for k, x in d["roofline_all"].items()})
PY
