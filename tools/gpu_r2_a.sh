#!/bin/bash
# round 2, run A: full GPU test suite after the refactor + baseline bench lines (cfg2, cfg3, cfg1)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -q -m gpu --timeout=300 -x --durations=8 > gpurun_out/r2a_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -25 gpurun_out/r2a_gpu_tests.log | cut -c1-220
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2a_bench_cfg2.json
timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-400 gpurun_out/r2a_bench_cfg3.json
timeout 300 python bench.py --config cfg1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_cfg1.json 2> gpurun_out/r2a_bench_cfg1.err; echo "cfg1 rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg1.json
