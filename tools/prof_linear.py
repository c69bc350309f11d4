"""Micro-driver for ncu: the cfg2 QKV projection (M=204800, N=192, K=64) through the tcgen05 Linear kernel."""
import sys
import torch
sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L
torch.manual_seed(0)
M, N, K = 204800, 192, 64
x = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda") * 0.2
b = torch.randn(N, device="cuda")
with torch.no_grad():
    for _ in range(4):
        y = ops.linear(x, w, b)
torch.cuda.synchronize()
print("ok")
