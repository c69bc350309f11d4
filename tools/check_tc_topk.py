"""tcgen05 full-catalogue top-k: exactness vs fp32 torch and throughput."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import rbm_b200
from rbm_b200 import ops

torch.manual_seed(0)
for (U, V, d, k, bias) in [(200, 20000, 64, 10, False), (1000, 100003, 64, 10, True), (300, 50000, 128, 10, False), (16384, 1000000, 64, 10, False)]:
    f = torch.randn(U, d, device="cuda")
    table = torch.randn(V + 1, d, device="cuda")
    b = torch.randn(V + 1, device="cuda") if bias else None
    vals, ids = ops.score_topk(f, table, b, 1, V + 1, k)
    torch.cuda.synchronize()
    nchk = min(U, 512)
    sc = f[:nchk].double() @ table[1:].double().t()
    if bias:
        sc = sc + b[1:].double()
    rv, ri = torch.topk(sc, k, dim=1)
    same = (ids[:nchk] == ri + 1).all(1).float().mean().item()
    err = (vals[:nchk].double() - rv).abs().max().item()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        ops.score_topk(f, table, b, 1, V + 1, k)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    print("U=%d V=%d d=%d bias=%s: rows with identical top-%d ids %.4f, max score err %.2e, %.2f ms -> %.0f users/s, %.1f TFLOP/s" % (
        U, V, d, bias, k, same, err, ms, U / ms * 1e3, 2.0 * U * V * d / ms / 1e9))
