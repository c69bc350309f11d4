import math, sys, torch
sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L
B, Ln, h = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dk = 32
d = h * dk
torch.manual_seed(0)
qkv = torch.randn(B * Ln, 3 * d, device="cuda")
tok = torch.ones(B, Ln, dtype=torch.long, device="cuda")
tok[0, : Ln // 3] = 0
out = ops.attention(qkv, None, tok, B, Ln, h, 0, d, 2 * d, mode, 1 / math.sqrt(dk), 0.0, 5, 11)
torch.cuda.synchronize()
q, k, v = (qkv[:, i * d:(i + 1) * d].double().view(B, Ln, h, dk).transpose(1, 2) for i in range(3))
s = q @ k.transpose(-1, -2) / math.sqrt(dk)
if mode == 2:
    s = s.masked_fill((tok == 0)[:, None, None, :], -1e9)
if mode == 1:
    s = s.masked_fill(~torch.tril(torch.ones(Ln, Ln, dtype=torch.bool, device="cuda")), float("-inf"))
ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * Ln, d)
print(sys.argv[1:], "err", (out.double() - ref).abs().max().item())
