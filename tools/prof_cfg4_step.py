#!/usr/bin/env python
"""Per-entry-point device times of one BERT4Rec configs[3] training step (d=256, nb=4, h=4, L=200, V=10^6) on one GPU."""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbm_b200  # noqa: E402
from rbm_b200 import lib as L  # noqa: E402

dev = "cuda"
V, Ln, d, B = 1_000_000, 200, 256, int(os.environ.get("B", "512"))
a = SimpleNamespace(model_code="bert", num_items=V, max_len=Ln, device=dev, model_init_seed=0, bert_num_blocks=4, bert_num_heads=4,
                    bert_hidden_units=d, bert_dropout=0.1, bert_hidden_dropout=0.1, optimizer="Adam", lr=1e-3, weight_decay=0, momentum=None,
                    decay_step=25, gamma=1.0, num_epochs=1, metric_ks=[10], best_metric="NDCG@10", train_batch_size=B, resume_path=None)
with torch.device(dev):
    m = rbm_b200.model_factory(a)
t = rbm_b200.trainer_factory(a, m, None, None, None, None)
m.train()
g = torch.Generator(device=dev).manual_seed(4)
tok = torch.randint(1, V + 1, (B, Ln), device=dev, generator=g)
lab = torch.where(torch.rand(B, Ln, device=dev, generator=g) < 0.15, tok, torch.zeros_like(tok))
batch = (torch.where(lab != 0, torch.full_like(tok, V + 1), tok), lab)
for _ in range(2):
    t.train_step(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    t.train_step(batch)
e1.record()
torch.cuda.synchronize()
print("step %.2f ms" % (e0.elapsed_time(e1) / 3))
L.profile = {}
t.train_step(batch)
prof = L.profile_collect()
L.profile = None
tot = {k: sum(ms for ms, _ in v) for k, v in prof.items()}
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-32s %8.3f ms  %5.1f %%  (%d calls)" % (k, v, 100 * v / s, len(prof[k])))
print("sum of entry points %.2f ms" % s)
# Linear family by shape: rbm_linear_fwd args (.., M=7, N=8, K=9), bwd_data (M=5, N=6, K=7), bwd_weight (M=6, N=7, K=8)
import collections
by = collections.defaultdict(lambda: [0.0, 0])
for name, idx in (("rbm_linear_fwd", (7, 8, 9)), ("rbm_linear_bwd_data", (5, 6, 7)), ("rbm_linear_bwd_weight", (6, 7, 8))):
    for ms, a in prof.get(name, []):
        key = (name, a[idx[1]], a[idx[2]])
        by[key][0] += ms
        by[key][1] += 1
for (name, N, K), (ms, n) in sorted(by.items(), key=lambda kv: -kv[1][0]):
    M = B * Ln
    print("%-24s N=%-5d K=%-5d %7.3f ms / %d calls = %.3f ms each  (%.0f TFLOP/s)" % (name, N, K, ms, n, ms / n, 2.0 * M * N * K / (ms / n) / 1e9))
