#!/bin/bash
# round 2, run B: re-run the tests that failed in run A, timeline of the persistent per-atom stage, launch lists
mkdir -p gpurun_out
export SPK_B200_CHAIN=1 SPK_B200_CFCONV=tc
timeout 600 python -m pytest tests/test_cuda_neighbors.py tests/test_script.py tests/test_reference_live.py -q -m gpu --timeout=300 > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2b_tests.log | cut -c1-220
timeout 200 python tools/chain_trace.py 256 > gpurun_out/r2b_chain_trace.txt 2>&1; echo "trace rc=$?"; cat gpurun_out/r2b_chain_trace.txt | cut -c1-230
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches_cfg3.csv \
    python bench.py --config cfg3 --steps 2 --warmup 2 --no-cpu-baseline --no-spatial > gpurun_out/r2b_under_ncu_cfg3.log 2>&1; echo "cfg3 launchlist rc=$?"
python - <<'PY'
import csv, collections
rows=[l for l in open("gpurun_out/r2b_launches_cfg3.csv") if l.startswith('"')]
r=csv.reader(rows); hdr=next(r); ki,vi,ui=hdr.index("Kernel Name"),hdr.index("Metric Value"),hdr.index("Metric Unit")
agg=collections.OrderedDict(); tot=0
seq=[]
for row in r:
    v=float(row[vi].replace(",","")); v = v/1000 if row[ui]=="ns" else v*1000 if row[ui]=="ms" else v
    seq.append((row[ki][:60],v))
for n,v in seq[-90:]:
    agg.setdefault(n,[0,0.0]); agg[n][0]+=1; agg[n][1]+=v; tot+=v
print("cfg3: last 90 launches, total us", round(tot,1))
for n,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"  {n:60s} {c:3d} {v:9.1f} us  avg {v/c:8.1f}")
PY
timeout 300 python bench.py --config cfg5 --atoms 65536 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_cfg5_64k.json 2> gpurun_out/r2b_bench_cfg5_64k.err; echo "cfg5(64k) rc=$?"; cut -c1-500 gpurun_out/r2b_bench_cfg5_64k.json; tail -5 gpurun_out/r2b_bench_cfg5_64k.err
