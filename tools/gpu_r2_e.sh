#!/bin/bash
# round 2, run E (1 GPU): accumulator-interval variant of k_dense_tc (parity + bench), phase trace of the fused SchNet kernel,
# cfg5 at full size on one GPU (strong-scaling baseline, tables > L2) with ncu DRAM bytes of its edge kernels, cfg2 launch list
mkdir -p gpurun_out
V=schnetpack_b200/csrc/libspk_b200_acc2.so
SPK_B200_LIB=$V timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_kernels.py -q -m gpu --timeout=300 -k "golden or cfg or dense or lin" > gpurun_out/r2e_acc2_tests.log 2>&1; echo "acc2 tests rc=$?"; tail -5 gpurun_out/r2e_acc2_tests.log | cut -c1-200
for lib in schnetpack_b200/csrc/libspk_b200.so $V; do
  SPK_B200_LIB=$lib timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-spatial > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench [$lib] rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench.json')); print(round(d['ms_per_step'],4), round(d['e2e']['value']), d['gpu_launches'])"
done
SPK_B200_LIB=schnetpack_b200/csrc/libspk_b200_strace.so timeout 200 python tools/schnet_trace.py > gpurun_out/r2e_schnet_trace.txt 2>&1; echo "schnet trace rc=$?"; tail -26 gpurun_out/r2e_schnet_trace.txt | cut -c1-200
timeout 600 python bench.py --config cfg5 --steps 6 --warmup 3 > gpurun_out/r2e_cfg5_n1.json 2> gpurun_out/r2e_cfg5_n1.err; echo "cfg5 N=1 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2e_cfg5_n1.json')); print(round(d['ms_per_step'],2), 'ms/step', d['rank0'], {k:(round(v['avg_us'],1), round(v['frac'],3), round(v['share_of_step'],3)) for k,v in d['roofline_all'].items()}, d['cpu_baseline'])"; tail -3 gpurun_out/r2e_cfg5_n1.err
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
timeout 900 ncu --metrics $M --clock-control none -k regex:"k_painn_edge" -s 6 -c 6 --csv --log-file gpurun_out/r2_ncu_cfg5_edge_kernels.csv \
    python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_under_ncu_cfg5.log 2>&1; echo "cfg5 ncu rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-spatial > gpurun_out/r2e_under_ncu1.log 2>&1; echo "launchlist rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_cfg2.csv "r2 launch list, cfg2 (aspirin x256, PaiNN 128x3, E+F), default pipeline, eager launches, one timed step" > gpurun_out/r2_launch_summary_cfg2.csv 2> gpurun_out/r2e_summ.err; head -20 gpurun_out/r2_launch_summary_cfg2.csv
