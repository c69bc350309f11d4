"""Micro-driver for ncu: fused full-catalogue scoring + top-10 at U=16384, V=1M, d=64 (tcgen05 path)."""
import sys
import torch
sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops
torch.manual_seed(0)
U, V, d = 16384, 1_000_000, 64
f = torch.randn(U, d, device="cuda")
table = torch.randn(V + 1, d, device="cuda")
for _ in range(3):
    vals, ids = ops.score_topk(f, table, None, 1, V + 1, 10)
torch.cuda.synchronize()
print("ok", ids[0].tolist())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    vals, ids = ops.score_topk(f, table, None, 1, V + 1, 10)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("score_topk U=%d V=%d d=%d: %.2f ms = %.1f TFLOP/s logical" % (U, V, d, ms, 2.0 * U * V * d / ms / 1e9))
