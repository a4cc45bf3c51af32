#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_kernels.py -q -m gpu -x --timeout=120 -k "tcgen05" > gpurun_out/tc.log 2>&1
echo "tc rc=$?"; tail -4 gpurun_out/tc.log
SPK_B200_DENSE=tc timeout 900 python -m pytest tests/test_cuda_parity.py -q -m gpu -s --timeout=600 > gpurun_out/parity_tc.log 2>&1
echo "parity_tc rc=$?"
grep -E "^\.?(painn|schnet|cfg)|passed|failed" gpurun_out/parity_tc.log | cut -c1-330
SPK_B200_DENSE=tc timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
SPK_B200_DENSE=ffma timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ldg.json 2> gpurun_out/bench_ldg.err
python - <<'PY'
import json
for n in ("tc","ldg"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:(round(v["avg_us"],1), round(v["frac"],3)) for k,v in d["roofline_all"].items()})
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-800:])
PY
