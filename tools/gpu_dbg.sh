#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/dbg.py <<'PY'
import torch, math, sys
sys.path.insert(0, '.')
from schnetpack_b200 import ops
torch.manual_seed(3)
for (M,K,N) in [(77,256,128),(77,384,128),(300,384,64),(5376,384,128)]:
    A = torch.randn(M, K, device='cuda'); W = torch.randn(N, K, device='cuda')/math.sqrt(K)
    wp = ops.tc_pack_weight(W)
    Y = ops.dense_tc(A, wp, N)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t()
    print(M,K,N, float((Y.double()-ref).abs().max()/ref.abs().max()))
PY
timeout 300 compute-sanitizer --tool memcheck python /tmp/dbg.py > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?"
grep -v "^$" gpurun_out/sanitizer.log | head -60
