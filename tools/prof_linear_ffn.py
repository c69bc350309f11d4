"""Micro-driver for ncu: FFN-1 of cfg2 (M=204800, N=256, K=64, bias + tanh-GELU + dropout, pre-activation saved)."""
import sys
import torch
sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L
torch.manual_seed(0)
M, N, K = 204800, 256, 64
x = torch.randn(M, K, device="cuda", requires_grad=True)
w = (torch.randn(N, K, device="cuda") * 0.2).requires_grad_(True)
b = torch.randn(N, device="cuda", requires_grad=True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(4):
    ev[0].record()
    y = ops.linear(x, w, b, act=L.ACT_GELU_TANH, pA=0.1, siteA=3, seed=5)
    ev[1].record()
torch.cuda.synchronize()
print("FFN-1 fwd %.3f ms" % ev[0].elapsed_time(ev[1]))
