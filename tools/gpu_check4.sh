#!/bin/bash
# full check: kernels (both dense impls), parity (default + tc), stress of tc tests, bench (default + tc), smoke
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_cuda_kernels.py -q -m gpu --timeout=300 > gpurun_out/kernels.log 2>&1; echo "kernels rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/kernels.log
timeout 900 python -m pytest tests/test_cuda_parity.py -q -m gpu --timeout=600 > gpurun_out/parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/parity.log | cut -c1-200
SPK_B200_DENSE=tc timeout 900 python -m pytest tests/test_cuda_parity.py -q -m gpu --timeout=600 > gpurun_out/parity_tc.log 2>&1; echo "parity_tc rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/parity_tc.log
for i in 1 2 3 4 5; do timeout 300 python -m pytest tests/test_cuda_kernels.py -q -m gpu --timeout=120 -k tcgen05 > gpurun_out/tc_stress_$i.log 2>&1; echo "tc stress $i rc=$?" | tee -a gpurun_out/summary.txt; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
SPK_B200_DENSE=tc timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench_tc rc=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
for n in ("bench","bench_tc"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("ms_per_step_median"), d["e2e"].get("api"), {k:(round(v["avg_us"],1), round(v["frac"],3)) for k,v in d["roofline_all"].items()})
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
