#!/bin/bash
# A/B timing of library variants: tools/gpu_r2_ab.sh [--test] <variant names...>   ("" = default build is always run first)
mkdir -p gpurun_out
C=schnetpack_b200/csrc
if [ "$1" == "--test" ]; then shift
  timeout 900 python -m pytest tests/test_cuda_kernels.py tests/test_cuda_parity.py -q -m gpu -x --timeout=300 2>&1 | tail -3
fi
for v in "" $@; do
  lib=$C/libspk_b200${v:+_$v}.so
  for cfg in cfg2 cfg3 cfg4; do
    SPK_B200_LIB=$lib timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline --no-spatial > gpurun_out/ab_${v:-base}_$cfg.json 2> gpurun_out/ab_${v:-base}_$cfg.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${v:-base}_$cfg.json"))
    print("${v:-base}", "$cfg", "ms/step", round(d["ms_per_step"], 4), {k: round(x["avg_us"], 1) for k, x in (d.get("roofline_all") or {}).items()})
except Exception as e:
    print("${v:-base} $cfg failed", e); print(open("gpurun_out/ab_${v:-base}_$cfg.err").read()[-800:])
PY
  done
done
