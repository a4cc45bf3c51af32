#!/bin/bash
# final-state profiles for profiles/: launch list (1 step), full-set of edge + dense kernels, other configs
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launchlist rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_painn_edge|k_dense_tc" -s 20 -c 10 -f -o gpurun_out/prof_final \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu2.log 2>&1
echo "fullset rc=$?"
for c in cfg4 cfg1 cfg3; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  echo "$c rc=$?"; python - $c <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), "edge-msgs/s", f'{d["edge_msgs_per_s"]:.3e}', "e2e", round(d["e2e"]["value"],1), {k:(round(v["avg_us"],1), round(v["frac"],3)) for k,v in (d["roofline_all"] or {}).items()}, d["config"]["atoms"], d["config"]["edges"])
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/bench_{sys.argv[1]}.err").read()[-600:])
PY
done
