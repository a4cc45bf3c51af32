#!/bin/bash
# ncu passes: (1) launch list with durations for ~2 steps, (2) full-set capture of the fused edge kernels.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launchlist rc=$?" | tee -a gpurun_out/summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_painn_edge -s 12 -c 6 -f -o gpurun_out/prof_edge \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
echo "fullset rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out | tail -12
