"""Development aid: per-chunk phase stamps (clock64, CTA 0) of k_schnet_cfconv_fwd_tc from a -DSPK_SCHNET_TRACE build.
usage (GPU box): SPK_B200_LIB=schnetpack_b200/csrc/libspk_b200_strace.so python tools/schnet_trace.py"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from schnetpack_b200 import _lib, ops  # noqa: E402
from schnetpack_b200 import synthetic as S  # noqa: E402
from schnetpack_b200.model import batch_to_device, from_spec  # noqa: E402

ops.CFCONV_IMPL = "tc"
dev = torch.device("cuda:0")
spec, data = S.make_config("cfg3")
model = from_spec(spec, S.init_params(spec, 0), dev)
x = batch_to_device(data, dev)
for _ in range(3):
    y = model.output_modules[0](model.representation(dict(x)))
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (256 * 16))()
h = _lib.lib()
h.spk_debug_schnet_trace.argtypes = [ctypes.c_void_p]
assert h.spk_debug_schnet_trace(buf) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(256, 16).astype(np.float64)
names = ["prod_pub", "mma1_iss", "b2_avail", "d_free", "mma2_iss", "act_Hrdy", "act_comp", "act_b2free", "act_stored", "cons_Drdy",
         "cons_done"]
t0 = t[0, 0]
print("chunk " + " ".join(f"{n:>10s}" for n in names))
for k in range(4, 24):
    print(f"{k:5d} " + " ".join(f"{t[k, i] - t0:10.0f}" for i in range(len(names))))
per = np.diff(t[8:60, 4])
print("period of MMA-2 issue (cycles): mean", per.mean(), "min", per.min(), "max", per.max())
