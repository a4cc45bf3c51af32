#!/bin/bash
# round 2 final validation on one GPU: full GPU suite, smoke, the driver's bench command (+ reference arm), the other
# configs, ncu launch list + metrics of the dominant kernels for profiles/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout=400 --durations=5 > gpurun_out/r2z_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/r2z_gpu_tests.log | cut -c1-160
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d.get("roofline") or {}
    print("  ", d.get("impl", "ours"), d["config"].get("workload", "")[:40], "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"], 1), d["unit"],
          "e2e", round(d["e2e"]["value"], 1), "roofline", r.get("kernel"), round(r.get("avg_us") or 0, 1), round(r.get("frac") or 0, 3),
          "clocks", d.get("clocks"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "eager", d.get("eager_gpu_baseline"))
    sp = d.get("spatial")
    if sp: print("   spatial", round(sp["ms_per_step"], 2), "ms", {k: (round(v["avg_us"], 1), round(v["frac"], 3)) for k, v in sp["roofline_all"].items()})
except Exception as e:
    print("   failed:", e)
PY
}
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; line gpurun_out/r2z_bench.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "ref arm rc=$?"; line gpurun_out/r2z_bench_ref.json
for cfg in cfg3 cfg1 cfg4; do
  timeout 400 python bench.py --config $cfg --steps 30 --warmup 5 --no-spatial > gpurun_out/r2z_bench_$cfg.json 2> gpurun_out/r2z_bench_$cfg.err; echo "$cfg rc=$?"; line gpurun_out/r2z_bench_$cfg.json
done
timeout 400 python bench.py --config cfg4 --md --steps 200 --warmup 5 --no-spatial --no-cpu-baseline > gpurun_out/r2z_bench_cfg4_md.json 2> gpurun_out/r2z_bench_cfg4_md.err; echo "cfg4 md rc=$?"; line gpurun_out/r2z_bench_cfg4_md.json
timeout 400 python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_cfg5.json 2> gpurun_out/r2z_bench_cfg5.err; echo "cfg5 rc=$?"; line gpurun_out/r2z_bench_cfg5.json
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_launches_cfg2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-spatial > gpurun_out/r2z_under_ncu1.log 2>&1; echo "launchlist rc=$?"
python tools/summarize_launches.py gpurun_out/r2z_launches_cfg2.csv "r2 FINAL launch list, cfg2 (aspirin x256, PaiNN 128x3, E+F), eager launches, one timed step" > gpurun_out/r2z_launch_summary_cfg2.csv 2>gpurun_out/r2z_summ.err; head -12 gpurun_out/r2z_launch_summary_cfg2.csv
timeout 500 ncu --metrics $M --clock-control none -k regex:"k_painn_edge|k_dense_tc" -s 40 -c 40 --csv --log-file gpurun_out/r2z_ncu_cfg2_kernels.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-spatial > gpurun_out/r2z_under_ncu2.log 2>&1; echo "cfg2 metrics rc=$?"
timeout 500 ncu --metrics $M --clock-control none -k regex:"k_schnet_cfconv_fwd_tc" -s 6 -c 6 --csv --log-file gpurun_out/r2z_ncu_cfg3_kernels.csv \
    python bench.py --config cfg3 --steps 2 --warmup 2 --no-cpu-baseline --no-spatial > gpurun_out/r2z_under_ncu3.log 2>&1; echo "cfg3 metrics rc=$?"
