"""Per-kernel timeline of one CUDA-graph replay INSIDE the programmatic-launch chain (debug build -DSPK_TIMELINE:
tools/build_variant.sh timeline -DSPK_TIMELINE; run with SPK_B200_LIB=schnetpack_b200/csrc/libspk_b200_timeline.so).
Every kernel stamps %globaltimer at entry and when its griddepcontrol.wait returns; consecutive wait-return stamps bracket
the dependent part of each kernel as it runs in the real chain.  usage: python tools/timeline.py [cfg2] [n_replays]"""
import ctypes, glob, os, re, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from schnetpack_b200 import _lib, synthetic as S
from schnetpack_b200.model import from_spec, batch_to_device, GraphedPotential

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
spec, data = S.make_config(cfg)
dev = torch.device("cuda:0")
model = from_spec(spec, S.init_params(spec, seed=0), dev)
gp = GraphedPotential(model)
batch = batch_to_device(data, dev)
gp(batch)
for _ in range(5):
    gp.replay()
torch.cuda.synchronize()
h = ctypes.CDLL(_lib.LIB_PATH)
CS = os.path.join(ROOT, "schnetpack_b200", "csrc")
tus = sorted(os.path.basename(f)[:-3] for f in glob.glob(os.path.join(CS, "*.cu")))


def dump():
    ev = []
    for tu in tus:
        fn = getattr(h, "spk_debug_timeline_" + tu, None)
        if fn is None:
            continue
        buf = np.zeros(2048, dtype=np.uint64)
        n = ctypes.c_uint(0)
        fn(buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n))
        for i in range(n.value):
            tag = int(buf[2 * i + 1])
            if tag >> 40:
                ev.append((int(buf[2 * i]), tu, 0, tag & 0xff))          # phase stamp (kind >= 2)
            else:
                ev.append((int(buf[2 * i]), tu, tag >> 1, tag & 1))
    return sorted(ev)


src = {}
def kernel_at(tu, line):
    """name of the __global__ function enclosing tu.cu:line (stamps come from macros in common.cuh's users; helper headers
    stamp with the line of the macro use inside the kernel body)"""
    if tu not in src:
        src[tu] = open(os.path.join(CS, tu + ".cu")).read().splitlines()
    for l in range(min(line, len(src[tu])) - 1, -1, -1):
        m = re.search(r"\b(k_[A-Za-z0-9_]+)\s*\(", src[tu][l])
        if m and ("__global__" in src[tu][l] or (l > 0 and "__global__" in src[tu][l - 1]) or "__launch_bounds__" in src[tu][l]):
            return m.group(1)
    return f"{tu}:{line}"


dump()
acc = {}
for r in range(reps):
    gp.replay()
    torch.cuda.synchronize()
    ev = dump()
    waits = [(t, tu, line) for t, tu, line, kind in ev if kind == 1]
    entries = [(t, tu, line) for t, tu, line, kind in ev if kind == 0]
    rows = []
    for i, (t, tu, line) in enumerate(waits):
        nxt = waits[i + 1][0] if i + 1 < len(waits) else None
        # entry stamp of the same kernel: the latest entry stamp of (tu, a line <= line) before t
        ent = [e for e in entries if e[1] == tu and e[0] <= t]
        lead = t - ent[-1][0] if ent else 0
        ph = {kind: pt - t for pt, ptu, _, kind in ev if kind >= 2 and ptu == tu and t <= pt < (nxt or t + 10**9)}
        rows.append((kernel_at(tu, line), (nxt - t) if nxt else 0, lead, ph))
    acc.setdefault(len(rows), []).append(rows)
nrows = max(acc, key=lambda k: len(acc[k]))
runs = acc[nrows]
print(f"# {cfg}: {nrows} kernels per replay, median over {len(runs)} replays; dur = wait-return(k+1) - wait-return(k) in us,")
print("# lead = wait-return - entry of the same kernel (how long it sat resident before its predecessor finished)")
tot = 0.0
agg = {}
for i in range(nrows):
    name = runs[0][i][0]
    dur = float(np.median([r[i][1] for r in runs])) / 1e3
    lead = float(np.median([r[i][2] for r in runs])) / 1e3
    tot += dur
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += dur
    kinds = sorted(set().union(*[r[i][3].keys() for r in runs]))
    ph = "  ".join(f"p{k}@{float(np.median([r[i][3].get(k, 0) for r in runs])) / 1e3:.2f}" for k in kinds)
    print(f"{i:3d} {name:28s} dur {dur:7.2f}  lead {lead:7.2f}  {ph}")
print(f"# sum {tot:.1f} us")
print("# per kernel: launches, total us, share")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"# {name:28s} {n:3d} {t:8.1f} {100 * t / tot:5.1f}%")
