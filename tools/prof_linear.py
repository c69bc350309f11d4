"""Micro-driver for ncu: the cfg2 QKV projection (M=204800, N=192, K=64) forward + backward through the tcgen05 Linear
kernels (forward, backward-data via the transposed weight, backward-weight with fused bias gradient)."""
import sys
import torch
sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L
torch.manual_seed(0)
M, N, K = 204800, 192, 64
x = torch.randn(M, K, device="cuda", requires_grad=True)
w = (torch.randn(N, K, device="cuda") * 0.2).requires_grad_(True)
b = torch.randn(N, device="cuda", requires_grad=True)
dy = torch.randn(M, N, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for _ in range(4):
    ev[0].record()
    y = ops.linear(x, w, b)
    ev[1].record()
    y.backward(dy)
    ev[2].record()
torch.cuda.synchronize()
print("linear M=%d N=%d K=%d  fwd %.3f ms  bwd (data + weight) %.3f ms" % (M, N, K, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
