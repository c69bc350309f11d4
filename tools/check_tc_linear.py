"""tcgen05 Linear bring-up: error vs an fp64 reference and CUDA-event time, for the cfg2 shapes."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L

torch.manual_seed(0)
for (M, N, K) in [(300, 192, 64), (257, 64, 256), (204800, 192, 64), (204800, 64, 64), (204800, 256, 64), (204800, 64, 256)]:
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.2
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda")
    xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y = ops.linear(xg, wg, b, residual=r, act=L.ACT_GELU_TANH)
    ref = torch.nn.functional.gelu(x.double() @ w.double().t() + b.double(), approximate="tanh") + r.double()
    err = (y.double() - ref).abs().max().item()
    dy = torch.randn(M, N, device="cuda")
    y.backward(dy)
    # reference grads in fp64
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    (torch.nn.functional.gelu(xd @ wd.t() + b.double(), approximate="tanh") + r.double()).backward(dy.double())
    ex = (xg.grad.double() - xd.grad).abs().max().item() / xd.grad.abs().max().item()
    ew = (wg.grad.double() - wd.grad).abs().max().item() / wd.grad.abs().max().item()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.no_grad():
        for _ in range(3):
            ops.linear(x, w, b, residual=r, act=L.ACT_GELU_TANH)
        ev[0].record()
        for _ in range(10):
            ops.linear(x, w, b, residual=r, act=L.ACT_GELU_TANH)
        ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    print("M=%d N=%d K=%d  fwd max|err|=%.2e (scale %.1f)  dx rel=%.2e dw rel=%.2e  fwd %.3f ms = %.1f TFLOP/s, %.0f GB/s" % (
        M, N, K, err, ref.abs().max().item(), ex, ew, ms, 2.0 * M * N * K / ms / 1e9, (M * K + 2 * M * N + N * K) * 4 / ms / 1e6))
