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
