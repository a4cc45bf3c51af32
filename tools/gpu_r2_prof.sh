#!/bin/bash
# round 2 profiling run (1 GPU): ncu launch list of one cfg2 step (eager launches so every kernel is its own node), full-set
# captures of the kernels this round added or that the roofline line names, exported to small CSVs on the box.
# usage: bash tools/gpu_r2_prof.sh   (env SPK_B200_CHAIN / SPK_B200_CFCONV select the pipelines)
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_under_ncu1.log 2>&1; echo "launchlist rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_cfg2.csv "r2 launch list, cfg2 (aspirin x256, PaiNN 128x3, E+F), atom chain=$SPK_B200_CHAIN, eager launches, one timed step" > gpurun_out/r2_launch_summary_cfg2.csv 2>gpurun_out/r2_summ.err; head -30 gpurun_out/r2_launch_summary_cfg2.csv
timeout 500 ncu --metrics $M --clock-control none -k regex:"k_atom_chain|k_painn_edge" -s 8 -c 24 --csv --log-file gpurun_out/r2_ncu_cfg2_kernels.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_under_ncu2.log 2>&1; echo "cfg2 metrics rc=$?"
timeout 500 ncu --metrics $M --clock-control none -k regex:"k_schnet_cfconv_fwd_tc|k_atom_chain" -s 6 -c 12 --csv --log-file gpurun_out/r2_ncu_cfg3_kernels.csv \
    python bench.py --config cfg3 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_under_ncu3.log 2>&1; echo "cfg3 metrics rc=$?"
timeout 700 ncu --metrics $M --clock-control none -k regex:"k_painn_edge" -s 6 -c 6 --csv --log-file gpurun_out/r2_ncu_cfg5_edge_kernels.csv \
    python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_under_ncu4.log 2>&1; echo "cfg5 metrics rc=$?"
ls -la gpurun_out/*.csv
