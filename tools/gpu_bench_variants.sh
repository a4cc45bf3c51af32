#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_kernels.py -q -m gpu --timeout=300 > gpurun_out/kernels.log 2>&1; echo "kernels rc=$?"; tail -2 gpurun_out/kernels.log
for v in "ffma ldg" "tc ldg" "ffma tma" "ffma sys"; do
  set -- $v
  SPK_B200_DENSE=$1 SPK_B200_EDGE=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
n = f"gpurun_out/bench_{sys.argv[1]}_{sys.argv[2]}"
try:
    d = json.load(open(n + ".json"))
    print(sys.argv[1:], "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:(round(v["avg_us"],1), round(v["frac"],3)) for k,v in d["roofline_all"].items()})
except Exception as e:
    print(sys.argv[1:], "failed", e); print(open(n + ".err").read()[-800:])
PY
done
