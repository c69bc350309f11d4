"""tcgen05 attention backward bring-up: per-block (dq/dk/dv x head) error against an fp64 autograd reference."""
import math
import sys

import torch

sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L

torch.manual_seed(0)
cases = [(2, 16, 1, L.MASK_NONE, 0.0), (2, 200, 2, L.MASK_KEYPAD, 0.0), (2, 200, 2, L.MASK_CAUSAL, 0.0), (3, 100, 2, L.MASK_KEYPAD, 0.2),
         (2, 256, 1, L.MASK_KEYPAD, 0.1), (5, 200, 2, L.MASK_KEYPAD, 0.1)]
if len(sys.argv) > 1:
    cases = cases[int(sys.argv[1]):int(sys.argv[1]) + 1]
for (B, Ln, h, mode, p) in cases:
    dk = 32
    d = h * dk
    qkv = torch.randn(B * Ln, 3 * d, device="cuda", requires_grad=True)
    tok = torch.randint(1, 50, (B, Ln), device="cuda")
    tok[0, : Ln // 3] = 0
    scale = 1 / math.sqrt(dk)
    out = ops.attention(qkv, None, tok, B, Ln, h, 0, d, 2 * d, mode, scale, p, 5, 11)
    dout = torch.randn_like(out)
    out.backward(dout)
    torch.cuda.synchronize()
    qd = qkv.detach().double().requires_grad_(True)
    q, k, v = (qd[:, i * d:(i + 1) * d].view(B, Ln, h, dk).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) * scale
    if mode == L.MASK_CAUSAL:
        s = s.masked_fill(~torch.tril(torch.ones(Ln, Ln, dtype=torch.bool, device="cuda")), float("-inf"))
    elif mode == L.MASK_KEYPAD:
        s = s.masked_fill((tok == 0)[:, None, None, :], -1e9)
    pr = torch.softmax(s, -1)
    if p > 0:
        mask = ops.dropout_mask_attn(B * h * Ln, Ln, p, 5, 11, "cuda").view(B, h, Ln, Ln).double()
        pr = pr * mask / (1 - p)
    ref = (pr @ v).transpose(1, 2).reshape(B * Ln, d)
    ref.backward(dout.double())
    g, gr = qkv.grad.double(), qd.grad
    msg = []
    for bi, name in enumerate(["dq", "dk", "dv"]):
        for hh in range(h):
            sl = slice(bi * d + hh * dk, bi * d + (hh + 1) * dk)
            e = (g[:, sl] - gr[:, sl]).abs().max().item() / max(gr[:, sl].abs().max().item(), 1e-30)
            msg.append("%s.h%d %.1e" % (name, hh, e))
    print("B=%d L=%d h=%d mode=%d p=%.1f  " % (B, Ln, h, mode, p) + "  ".join(msg))
