#!/bin/bash
mkdir -p gpurun_out
SPK_B200_DENSE=tc timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 300 --csv --log-file gpurun_out/launches_tc.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_tc.log 2>&1
echo "launchlist rc=$?"
SPK_B200_DENSE=tc timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_tc -s 40 -c 4 -f -o gpurun_out/prof_tc \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_tc2.log 2>&1
echo "fullset rc=$?"
