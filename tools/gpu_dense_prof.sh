#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense -s 30 -c 6 -f -o gpurun_out/prof_dense \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu3.log 2>&1
echo "fullset rc=$?"
