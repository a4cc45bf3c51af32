#!/bin/bash
# round 2, run A: GPU test suite after the refactor + bench lines (cfg2, cfg3, cfg1); the new persistent per-atom stage
# is smoke-tested first under a short timeout and switched off for the rest of the run if it fails
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python -m pytest tests/test_cuda_kernels.py -q -m gpu -k "atom_chain" --timeout=120 -x > gpurun_out/r2a_chain.log 2>&1; rc=$?
echo "chain smoke rc=$rc"; tail -15 gpurun_out/r2a_chain.log | cut -c1-200
if [ $rc -ne 0 ]; then export SPK_B200_CHAIN=0; echo "!! atom chain OFF for the rest of this run"; fi
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 --durations=8 > gpurun_out/r2a_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -40 gpurun_out/r2a_gpu_tests.log | cut -c1-220
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg2.json; tail -3 gpurun_out/r2a_bench_cfg2.err
SPK_B200_CHAIN=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_cfg2_nochain.json 2> gpurun_out/r2a_bench_cfg2_nochain.err; echo "bench(no chain) rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg2_nochain.json
timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg3.json; tail -3 gpurun_out/r2a_bench_cfg3.err
timeout 300 python bench.py --config cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_cfg1.json 2> gpurun_out/r2a_bench_cfg1.err; echo "cfg1 rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg1.json
timeout 400 python bench.py --config cfg5 --atoms 65536 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_cfg5_64k.json 2> gpurun_out/r2a_bench_cfg5_64k.err; echo "cfg5(64k) rc=$?"; cut -c1-400 gpurun_out/r2a_bench_cfg5_64k.json; tail -5 gpurun_out/r2a_bench_cfg5_64k.err
