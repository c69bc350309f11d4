"""CUDA-event timing of the tcgen05 Linear forward for the cfg2 shapes, per epilogue variant."""
import sys
import torch
sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L

torch.manual_seed(0)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(204800, 192, 64), (204800, 64, 64), (204800, 256, 64), (204800, 64, 256)]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.2
    b = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda")
    with torch.no_grad():
        variants = {
            "plain": lambda: ops.linear(x, w, b),
            "res": lambda: ops.linear(x, w, b, residual=r),
            "gelu": lambda: ops.linear(x, w, b, act=L.ACT_GELU_TANH),
            "drop+res": lambda: ops.linear(x, w, b, residual=r, pA=0.1, siteA=3, seed=5),
        }
        for name, fn in variants.items():
            ms = t(fn)
            byt = (M * K + M * N + N * K + (M * N if "res" in name else 0) + (M * N if name == "gelu" else 0)) * 4
            print("M=%d N=%d K=%d %-9s %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (M, N, K, name, ms, 2.0 * M * N * K / ms / 1e9, byt / ms / 1e6))
