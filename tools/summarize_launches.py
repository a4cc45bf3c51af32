"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares of ONE step.
usage: python tools/summarize_launches.py gpurun_out/launches.csv "<header comment>" > profiles/<name>.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for row in r:
    v = float(row[vi].replace(",", ""))
    v = v / 1000 if row[ui] == "ns" else v * 1000 if row[ui] == "ms" else v
    seq.append((row[ki], v))
starts = [i for i, (n, _) in enumerate(seq) if "k_init_status" in n]      # first kernel of spk_graph_build
a, b = starts[1], starts[2]
step = [(n, v) for n, v in seq[a:b] if "FillFunctor<unsigned char>" not in n]     # drop the bench's L2 flush
tot = sum(v for _, v in step)
agg = collections.OrderedDict()
for n, v in step:
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("<unnamed>::", "")
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += v
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# source: ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES")
print(f"# launches in step: {len(step)}, sum of kernel durations: {tot:.1f} us")
print("kernel,launches,total_us,share_pct,avg_us")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'"{n}",{c},{v:.1f},{100 * v / tot:.1f},{v / c:.1f}')
