"""Phase trace of the tensor-core edge kernel (debug build with -DSPK_EDGE_TRACE, SPK_B200_LIB=tools/build/libspk_trace.so)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SPK_B200_EDGE"] = "tc"
from schnetpack_b200 import _lib, ops, synthetic as S
from schnetpack_b200.model import from_spec, batch_to_device

spec, data = S.make_config("cfg2")
params = S.init_params(spec, seed=0)
dev = torch.device("cuda:0")
model = from_spec(spec, params, dev)
x = batch_to_device(data, dev)
for _ in range(3):
    ops._GRAPH_CACHE.clear()
    y = dict(x); y[S.R] = y[S.R].detach()
    out = model(y)
torch.cuda.synchronize()
h = _lib.lib()
buf = np.zeros(148 * 256, dtype=np.int64)
rc = h.spk_debug_edge_trace(buf.ctypes.data_as(ctypes.c_void_p))
print("rc", rc)
bwd = os.environ.get("SPK_EDGE_TRACE_KIND", "fwd") == "bwd"   # library built with -DSPK_EDGE_TRACE=2
for cta in (0, 73, 147):
    t = buf[cta * 256:(cta + 1) * 256]
    t0 = t[0]
    print(f"cta {cta}: setup {t[1]-t0} loop_end {t[2]-t0} exit {t[3]-t0}")
    if bwd:
        print("  k: mma_issued | cons_wait_start ready half0_done half1_done group_barrier chunk_done")
        for k in range(16):
            print(f"  {k:2d}: {t[48+k]-t0:8d} | {t[64+k]-t0:8d} {t[80+k]-t0:8d} {t[112+k]-t0:8d} {t[128+k]-t0:8d} {t[144+k]-t0:8d} {t[96+k]-t0:8d}")
        continue
    print("  k: prod_issued prod_published | mma_issued | cons_wait_start acc_ready w_half0 half0_done w_half1 edges_done")
    for k in range(10):
        pi = t[16 + k // 3] - t0 if k % 3 == 0 else -1
        pp = t[32 + k // 3] - t0 if k % 3 == 0 else -1
        print(f"  {k:2d}: {pi:8d} {pp:8d} | {t[48+k]-t0:8d} | {t[64+k]-t0:8d} {t[80+k]-t0:8d} {t[112+k]-t0:8d} {t[128+k]-t0:8d} {t[144+k]-t0:8d} {t[96+k]-t0:8d}")
