#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_cuda_kernels.py -q -m gpu -x --timeout=300 > gpurun_out/kernels.log 2>&1
echo "kernels rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/kernels.log
timeout 900 python -m pytest tests/test_cuda_parity.py -q -m gpu -s --timeout=600 > gpurun_out/parity.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/summary.txt
grep -E "^\.?(painn|schnet|cfg)|passed|failed" gpurun_out/parity.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:(round(v["avg_us"],1), round(v["frac"],3)) for k,v in d["roofline_all"].items()}, d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench.err").read()[-1500:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launchlist rc=$?" | tee -a gpurun_out/summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_painn_edge -s 12 -c 6 -f -o gpurun_out/prof_edge \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
echo "fullset rc=$?" | tee -a gpurun_out/summary.txt
