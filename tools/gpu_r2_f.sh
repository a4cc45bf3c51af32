#!/bin/bash
# round 2, run F (1 GPU): fused SchNet kernel with the activation folded into the consumer warps, partitioned engine after
# the receiver-row change (ranks sharing the GPU over gloo), full GPU suite with the round's defaults
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_kernels.py -q -m gpu -k "schnet_fused" --timeout=120 -x > gpurun_out/r2f_cfconv.log 2>&1; echo "schnet fused rc=$?"; tail -3 gpurun_out/r2f_cfconv.log | cut -c1-200
tools/build_variant.sh strace -DSPK_SCHNET_TRACE > /dev/null 2>&1 || true
SPK_B200_LIB=schnetpack_b200/csrc/libspk_b200_strace.so timeout 200 python tools/schnet_trace.py > gpurun_out/r2f_schnet_trace.txt 2>&1; echo "schnet trace rc=$?"; tail -8 gpurun_out/r2f_schnet_trace.txt | cut -c1-200
timeout 300 python bench.py --config cfg3 --steps 20 --warmup 3 --no-spatial > gpurun_out/r2f_bench_cfg3.json 2> gpurun_out/r2f_bench_cfg3.err; echo "cfg3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_cfg3.json')); print(round(d['ms_per_step'],4), d['e2e']['ms_per_step_median'], d['roofline']['avg_us'], d['roofline']['achieved'], d['eager_gpu_baseline'], d['cpu_baseline'])"
timeout 1500 python -m pytest tests -q -m gpu --timeout=400 --durations=6 > gpurun_out/r2f_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -16 gpurun_out/r2f_gpu_tests.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
