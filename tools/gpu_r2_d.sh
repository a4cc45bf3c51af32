#!/bin/bash
# round 2, run D (2 GPUs): the partitioned CUDA engine over NCCL ranks with both halo transports, cfg5 strong scaling at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_cuda_partitioned.py tests/test_script.py -q -m gpu --timeout=400 > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2d_tests.log | cut -c1-220
for tr in peer nccl; do
  SPK_B200_HALO=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --config cfg5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_cfg5_n2_$tr.json 2> gpurun_out/r2d_cfg5_n2_$tr.err
  echo "cfg5 N=2 [$tr] rc=$?"; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2d_cfg5_n2_$tr.json')); print(round(d['ms_per_step'],2), 'ms/step', d['halo'], d['rank0'], {k:(round(v['avg_us'],1), round(v['frac'],3)) for k,v in d['roofline_all'].items()})
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2d_cfg5_n2_$tr.err').read()[-1500:])"
done
