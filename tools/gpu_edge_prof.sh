#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_painn_edge -s 12 -c 6 -f -o gpurun_out/prof_edge \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu2.log 2>&1
echo "fullset rc=$?"
