import math, sys, torch
sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops, lib as L
B, Ln, h, dk = 1, 16, 1, 32
d = h * dk
torch.manual_seed(0)
qkv = torch.randn(B * Ln, 3 * d, device="cuda")
tok = torch.ones(B, Ln, dtype=torch.long, device="cuda")
out = ops.attention(qkv, None, tok, B, Ln, h, 0, d, 2 * d, L.MASK_NONE, 1 / math.sqrt(dk), 0.0, 5, 11)
torch.cuda.synchronize()
q, k, v = (qkv[:, i * d:(i + 1) * d].double() for i in range(3))
ref = torch.softmax(q @ k.t() / math.sqrt(dk), -1) @ v
print("err", (out.double() - ref).abs().max().item())
print("out[0,:8]", out[0, :8].tolist())
print("out[5,:8]", out[5, :8].tolist())
print("ref[0,:4]", ref[0, :4].tolist())
s = (q @ k.t() / math.sqrt(dk))
print("S[0,:2] nat", s[0, :2].tolist(), " S*log2e:", (s[0, :2] * 1.4426950408889634).tolist(), "raw qk:", (q @ k.t())[0, :2].tolist())
print("rowmax log2", (s[0].max() * 1.4426950408889634).item(), "P[0,:2]", torch.softmax(s, -1)[0, :2].tolist())
