"""score_topk at the BASELINE configs[4] shape (16384 users x 10^7 items, d = 64): kernel time of the current plan."""
import sys
import torch
sys.path.insert(0, ".")
from rbm_b200 import ops
torch.manual_seed(0)
U, V, d = 16384, 10_000_000, 64
f = torch.randn(U, d, device="cuda")
table = torch.randn(V + 1, d, device="cuda")
for _ in range(2):
    vals, ids = ops.score_topk(f, table, None, 1, V + 1, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    vals, ids = ops.score_topk(f, table, None, 1, V + 1, 10)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
sc = f[:64].double() @ table[1:].double().t()
ri = torch.topk(sc, 10, dim=1).indices + 1
print("score_topk U=%d V=%d d=%d: %.2f ms = %.0f users/s, %.1f TFLOP/s logical; identical top-10 rows %.3f" % (
    U, V, d, ms, U / ms * 1e3, 2.0 * U * V * d / ms / 1e9, (ids[:64] == ri).all(1).float().mean().item()))
