#!/bin/bash
# round 2, run C: chain kernel after the staging / epilogue / glue fixes, fused SchNet kernel with N-folded MMAs
mkdir -p gpurun_out
export SPK_B200_CHAIN=1 SPK_B200_CFCONV=tc
timeout 600 python -m pytest tests/test_cuda_kernels.py tests/test_script.py -q -m gpu --timeout=300 -k "atom_chain or schnet_fused or scripted" > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2c_tests.log | cut -c1-220
timeout 200 python tools/chain_trace.py 256 > gpurun_out/r2c_chain_trace.txt 2>&1; echo "trace rc=$?"; cat gpurun_out/r2c_chain_trace.txt | cut -c1-230
for v in "CHAIN=1" "CHAIN=1 SPK_B200_CHAIN_NFOLD=1" "CHAIN=0"; do
  env SPK_B200_$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-spatial > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench [$v] rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench.json')); print(round(d['ms_per_step'],4), round(d['e2e']['value']), d['gpu_launches'], {k:round(v['avg_us'],1) for k,v in d['roofline_all'].items()})"
done
timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-spatial > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err; echo "cfg3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_cfg3.json')); print(round(d['ms_per_step'],4), d['e2e'], d['roofline'])"
