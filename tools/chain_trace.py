"""Development aid: per-item timeline of the persistent per-atom stage (csrc/atom_chain.cu) on one cfg2 evaluation.
usage (GPU box): SPK_B200_CHAIN=1 python tools/chain_trace.py [batch]  ->  per stage and step: items, mean / max of the
dependency wait, the K-loop (or glue) time and the epilogue + publish time, and the stage's wall time."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from schnetpack_b200 import ops  # noqa: E402
from schnetpack_b200 import synthetic as S  # noqa: E402
from schnetpack_b200.model import batch_to_device, from_spec  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ops.CHAIN_IMPL = True
dev = torch.device("cuda:0")
spec, data = S.make_config("cfg2", batch=batch)
model = from_spec(spec, S.init_params(spec, 0), dev)
x = batch_to_device(data, dev)
for _ in range(3):
    model(dict(x))
torch.cuda.synchronize()
ops.CHAIN_TRACE = []
model(dict(x))
torch.cuda.synchronize()
KIND = {0: "gemm", 1: "ctx", 2: "upd", 3: "upd_bwd", 4: "ctx_bwd"}
for li, (tr, prog) in enumerate(ops.CHAIN_TRACE):
    t = tr.cpu().numpy()
    t0 = t[:, 0].min()
    print(f"== stage launch {li}: {len(prog)} steps, {t.shape[0]} items, wall {1e-3 * (t[:, 3].max() - t0):.1f} us, "
          f"CTAs used {len(np.unique(t[:, 4]))}")
    for si, (kind, K, N, rpa) in enumerate(prog):
        m = t[:, 5] == si
        if not m.any():
            continue
        r = t[m]
        wait, comp, epi = 1e-3 * (r[:, 1] - r[:, 0]), 1e-3 * (np.maximum(r[:, 2], r[:, 1]) - r[:, 1]), 1e-3 * (r[:, 3] - np.maximum(r[:, 2], r[:, 1]))
        print(f"   step {si} {KIND[kind]:8s} K={K:3d} N={N:3d} x{rpa}: items {m.sum():4d}  claim@ {1e-3 * (r[:, 0].min() - t0):7.1f}..{1e-3 * (r[:, 0].max() - t0):7.1f} us  "
              f"wait {wait.mean():6.1f}/{wait.max():6.1f}  work {comp.mean():6.1f}/{comp.max():6.1f}  epilogue+publish {epi.mean():5.1f}/{epi.max():5.1f}  "
              f"done@ {1e-3 * (r[:, 3].max() - t0):7.1f} us")
