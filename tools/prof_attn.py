"""Micro-driver: attention fwd+bwd at the cfg2 shape (B=1024, L=200, h=2, dk=32, key-padding mask, p=0.1).
Prints CUDA-event times; run under ncu for the per-kernel profile."""
import math
import sys

import torch

sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
Ln, h, dk = 200, 2, 32
d = h * dk
torch.manual_seed(0)
qkv = torch.randn(B * Ln, 3 * d, device="cuda", requires_grad=True)
tok = torch.randint(1, 100, (B, Ln), device="cuda")
tok[:, :20] = 0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    ev[0].record()
    out = ops.attention(qkv, None, tok, B, Ln, h, 0, d, 2 * d, L.MASK_KEYPAD, 1 / math.sqrt(dk), p, 7, 3)
    ev[1].record()
    out.backward(torch.ones_like(out))
    ev[2].record()
torch.cuda.synchronize()
flop = 4.0 * Ln * Ln * dk * B * h
print("attn fwd %.3f ms (%.1f TFLOP/s)  bwd %.3f ms (%.1f TFLOP/s)  p=%.2f" % (
    ev[0].elapsed_time(ev[1]), flop / ev[0].elapsed_time(ev[1]) / 1e9, ev[1].elapsed_time(ev[2]), 2.5 * flop / ev[1].elapsed_time(ev[2]) / 1e9, p))
