#!/bin/bash
# final validation of the round: all GPU tests, the bench line, other configs, ncu launch list + full-set capture
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -q -m gpu --timeout=300 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"
for c in cfg4 cfg1; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
python - <<'PY'
import json
for c in ("cfg2", "cfg4", "cfg1"):
    try:
        d = json.load(open(f"gpurun_out/bench_{c}.json"))
        print(c, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
              {k: (round(v["avg_us"], 1), round(v["frac"], 3), round(v["share_of_step"], 3)) for k, v in (d["roofline_all"] or {}).items()},
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "eager", (d.get("eager_gpu_baseline") or {}).get("value"),
              "launches", d.get("gpu_launches"), d.get("clocks"))
    except Exception as e:
        print(c, "failed", e); print(open(f"gpurun_out/bench_{c}.err").read()[-800:])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launchlist rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_painn_edge|k_dense_tc" -s 6 -c 30 -f -o gpurun_out/prof_final \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu2.log 2>&1
echo "fullset rc=$?"; ls -la gpurun_out/prof_final.ncu-rep
