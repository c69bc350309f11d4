"""Per-entry-point device time of one SASRec training step at the BASELINE configs[2] shape (bench.py's extra)."""
import sys, os
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbm_b200
from rbm_b200 import lib as L
from rbm_b200.dataloaders import synthetic_interactions, sliding_window_partition, SasBatcher

dev = torch.device("cuda", 0)
V, Ln, d, B = 12101, 50, 128, 4096
a = SimpleNamespace(model_code="sas", num_items=V, max_len=Ln, device=str(dev), sas_hidden_units=d, sas_num_blocks=2, sas_heads=2,
                    sas_dropout=0.2, l2_emb=0.0, optimizer="Adam", lr=1e-3, weight_decay=0, momentum=None, decay_step=25, gamma=1.0,
                    num_epochs=1, metric_ks=[10], best_metric="NDCG@10", train_batch_size=B, resume_path=None)
hist = synthetic_interactions(22363, V, 8.9, 5, seed=1234)
ds = sliding_window_partition(hist, Ln, 0.3)
sb = SasBatcher(ds[0], V, Ln, seed=1)
batches = [tuple(torch.from_numpy(x).to(dev) for x in sb.batch(B)) for _ in range(2)]
model = rbm_b200.model_factory(a)
trainer = rbm_b200.trainer_factory(a, model, None, None, None, None)
model.train()
for i in range(3):
    trainer.train_step(batches[i % 2])
torch.cuda.synchronize()
L.profile = {}
for i in range(3):
    trainer.train_step(batches[i % 2])
prof = L.profile_collect()
L.profile = None
tot = {k: sum(ms for ms, _ in v) / 3 for k, v in prof.items()}
s = sum(tot.values())
print("sum of entry points: %.3f ms/step" % s)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%6.1f%%  %8.1f us  x%-3d %s" % (100 * v / s, v * 1e3, len(prof[k]) // 3, k))
nz = (batches[0][0] != 0).float().mean().item()
print("non-pad fraction of the batch: %.3f" % nz)
