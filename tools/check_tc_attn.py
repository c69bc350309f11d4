"""tcgen05 attention forward bring-up: compare against the mma.sync kernel / fp64 reference at d_k = 32."""
import math
import os
import sys

import torch

sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L

torch.manual_seed(0)
for (B, Ln, h, mode, p) in [(2, 16, 1, L.MASK_NONE, 0.0), (2, 128, 2, L.MASK_KEYPAD, 0.0), (3, 200, 2, L.MASK_KEYPAD, 0.0), (2, 200, 2, L.MASK_CAUSAL, 0.0),
                            (2, 50, 2, L.MASK_CAUSAL, 0.0), (2, 256, 1, L.MASK_KEYPAD, 0.0), (3, 200, 2, L.MASK_KEYPAD, 0.2)]:
    dk = 32
    d = h * dk
    qkv = torch.randn(B * Ln, 3 * d, device="cuda")
    tok = torch.randint(1, 50, (B, Ln), device="cuda")
    tok[0, : Ln // 3] = 0
    scale = 1 / math.sqrt(dk)
    out = ops.attention(qkv, None, tok, B, Ln, h, 0, d, 2 * d, mode, scale, p, 5, 11)
    torch.cuda.synchronize()
    q, k, v = (qkv[:, i * d:(i + 1) * d].double().view(B, Ln, h, dk).transpose(1, 2) for i in range(3))
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
    err = (out.double() - ref).abs().max().item()
    print("B=%d L=%d h=%d mode=%d p=%.1f  max|err|=%.3e (scale %.2f)" % (B, Ln, h, mode, p, err, ref.abs().max().item()))
