"""Micro-driver: fused scoring + masked cross-entropy (fwd + bwd) at the cfg2 shape (204800 positions, 15% masked,
V+1 = 3417 items, d = 64).  Prints CUDA-event times; run under ncu for per-kernel profiles."""
import sys
import torch
sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops

torch.manual_seed(0)
M, V1, d = 204800, 3417, 64
h = torch.randn(M, d, device="cuda", requires_grad=True)
w = (torch.randn(V1, d, device="cuda") * 0.1).requires_grad_(True)
b = torch.zeros(V1, device="cuda", requires_grad=True)
labels = torch.randint(1, V1, (M,), device="cuda")
labels[torch.rand(M, device="cuda") > 0.15] = 0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    ev[0].record()
    loss = ops.score_cross_entropy(h, labels, w, b)
    ev[1].record()
    loss.backward()
    ev[2].record()
torch.cuda.synchronize()
n = int((labels != 0).sum())
fl = 2.0 * n * V1 * d
print("masked rows %d  loss %.4f  fwd %.3f ms (%.1f TFLOP/s)  bwd %.3f ms (%.1f TFLOP/s)" % (
    n, loss.item(), ev[0].elapsed_time(ev[1]), fl / ev[0].elapsed_time(ev[1]) / 1e9, ev[1].elapsed_time(ev[2]), 2 * fl / ev[1].elapsed_time(ev[2]) / 1e9))
