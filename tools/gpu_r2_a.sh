#!/bin/bash
# round 2, run A: validate the new kernels first (persistent per-atom stage, fused SchNet forward), each under a short timeout;
# whatever passes is switched ON for the full suite and the benches of this run
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python -m pytest tests/test_cuda_kernels.py -q -m gpu -k "atom_chain" --timeout=120 -x > gpurun_out/r2a_chain.log 2>&1; rc=$?
echo "chain smoke rc=$rc"; tail -15 gpurun_out/r2a_chain.log | cut -c1-200
if [ $rc -eq 0 ]; then export SPK_B200_CHAIN=1; else export SPK_B200_CHAIN=0; echo "!! atom chain OFF"; fi
timeout 300 python -m pytest tests/test_cuda_kernels.py -q -m gpu -k "schnet_fused" --timeout=120 -x > gpurun_out/r2a_cfconv.log 2>&1; rc=$?
echo "schnet fused smoke rc=$rc"; tail -15 gpurun_out/r2a_cfconv.log | cut -c1-200
if [ $rc -eq 0 ]; then export SPK_B200_CFCONV=tc; else export SPK_B200_CFCONV=mat; echo "!! fused schnet OFF"; fi
echo "switches: CHAIN=$SPK_B200_CHAIN CFCONV=$SPK_B200_CFCONV" | tee gpurun_out/r2a_switches.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout=300 --durations=8 > gpurun_out/r2a_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -40 gpurun_out/r2a_gpu_tests.log | cut -c1-220
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 500 python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg2.json; tail -3 gpurun_out/r2a_bench_cfg2.err
SPK_B200_CHAIN_NFOLD=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-spatial > gpurun_out/r2a_bench_cfg2_nfold.json 2> gpurun_out/r2a_bench_cfg2_nfold.err; echo "bench(nfold) rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg2_nfold.json
SPK_B200_CHAIN=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-spatial > gpurun_out/r2a_bench_cfg2_nochain.json 2> gpurun_out/r2a_bench_cfg2_nochain.err; echo "bench(no chain) rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg2_nochain.json
timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-spatial > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg3.json; tail -3 gpurun_out/r2a_bench_cfg3.err
SPK_B200_CFCONV=mat timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-spatial > gpurun_out/r2a_bench_cfg3_mat.json 2> gpurun_out/r2a_bench_cfg3_mat.err; echo "cfg3(mat) rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg3_mat.json
timeout 300 python bench.py --config cfg1 --steps 50 --warmup 5 --no-cpu-baseline --no-spatial > gpurun_out/r2a_bench_cfg1.json 2> gpurun_out/r2a_bench_cfg1.err; echo "cfg1 rc=$?"; cut -c1-300 gpurun_out/r2a_bench_cfg1.json
