#!/usr/bin/env python
"""Bring-up / timing driver of the split-fp16 tcgen05 scoring + cross-entropy kernels (csrc/ce_wide.cu):
    python tools/dbg_ce_wide.py check      # loss / dH / dW / db against an fp64 torch evaluation, several shapes
    python tools/dbg_ce_wide.py time       # cfg4 shape (d=256, V=10^6), per-entry-point times
    python tools/dbg_ce_wide.py prof       # the same calls once, for ncu"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbm_b200  # noqa: E402,F401
from rbm_b200 import ops, lib as L  # noqa: E402

DEV = "cuda"


def ref64(h, labels, w, b):
    sel = labels != 0
    h64, w64, b64 = h[sel].double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(h64 @ w64.t() + b64, labels[sel])
    loss.backward()
    dh = torch.zeros_like(h, dtype=torch.float64)
    dh[sel] = h64.grad
    return loss.item(), dh, w64.grad, b64.grad


def check(n, V1, d, rate, seed, hs=0.5, ws=0.2):
    g = torch.Generator(device=DEV).manual_seed(seed)
    h = (torch.randn(n, d, device=DEV, generator=g) * hs).requires_grad_(True)
    w = (torch.randn(V1, d, device=DEV, generator=g) * ws).requires_grad_(True)
    b = (torch.randn(V1, device=DEV, generator=g) * 0.1).requires_grad_(True)
    labels = torch.where(torch.rand(n, device=DEV, generator=g) < rate, torch.randint(1, V1, (n,), device=DEV, generator=g),
                         torch.zeros(n, dtype=torch.long, device=DEV))
    labels[0] = V1 - 1
    loss = ops.score_cross_entropy(h, labels, w, b)
    loss.backward()
    torch.cuda.synchronize()
    rl, rdh, rdw, rdb = ref64(h.detach(), labels, w.detach(), b.detach())
    e = lambda a, r: float((a.double() - r).abs().max() / r.abs().max().clamp_min(1e-30))
    print("n=%d V1=%d d=%d P=%d  loss %.7f ref %.7f rel %.2e | dH %.2e dW %.2e db %.2e" % (
        n, V1, d, int((labels != 0).sum()), loss.item(), rl, abs(loss.item() - rl) / abs(rl), e(h.grad, rdh), e(w.grad, rdw), e(b.grad, rdb)), flush=True)


def timing(B=512, V1=1_000_001, d=256, Ln=200, prof=False):
    g = torch.Generator(device=DEV).manual_seed(1)
    n = B * Ln
    h = (torch.randn(n, d, device=DEV, generator=g) * 0.5).requires_grad_(True)
    w = (torch.randn(V1, d, device=DEV, generator=g) * 0.05).requires_grad_(True)
    b = torch.zeros(V1, device=DEV).requires_grad_(True)
    labels = torch.where(torch.rand(n, device=DEV, generator=g) < 0.15, torch.randint(1, V1, (n,), device=DEV, generator=g),
                         torch.zeros(n, dtype=torch.long, device=DEV))
    P = int((labels != 0).sum())
    for it in range(1 if prof else 3):
        L.profile = {}
        h.grad = w.grad = b.grad = None
        loss = ops.score_cross_entropy(h, labels, w, b)
        loss.backward()
        prof_ms = L.profile_collect()
        L.profile = None
        f = prof_ms["rbm_ce_fwd"][0][0]
        bw = prof_ms["rbm_ce_bwd"][0][0]
        fl = 2.0 * d * V1 * P
        print("B=%d P=%d V1=%d d=%d: fwd %.2f ms (%.0f TFLOP/s)  bwd %.2f ms (%.0f TFLOP/s algorithmic 4dVP)  loss %.5f  -> %.0f seq/s (CE only)" % (
            B, P, V1, d, f, fl / f / 1e9, bw, 2 * fl / bw / 1e9, loss.item(), B / (f + bw) * 1e3), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        check(640, 900, 256, 0.3, 1)
        check(640, 900, 128, 0.3, 2)
        check(3000, 5000, 256, 0.2, 3)
        check(1000, 3417, 128, 0.15, 4)
        check(40000, 20011, 256, 0.15, 5)
        check(333, 64, 128, 0.5, 6)
        check(2048, 70001, 256, 0.1, 7, hs=30.0, ws=1e-3)
    elif mode == "time":
        timing(int(os.environ.get("B", "512")))
        timing(128)
    else:
        timing(int(os.environ.get("B", "512")), prof=True)
