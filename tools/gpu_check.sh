#!/bin/bash
# One GPU-box session: kernel numerics, parity, smoke, bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_cuda_kernels.py -q -m gpu -x --timeout=300 > gpurun_out/kernels.log 2>&1
echo "kernels rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_cuda_parity.py -q -m gpu -s --timeout=600 > gpurun_out/parity.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/kernels.log; tail -30 gpurun_out/parity.log; cat gpurun_out/smoke.log | tail -5; cat gpurun_out/bench.json | cut -c1-1500; tail -5 gpurun_out/bench.err
