import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from types import SimpleNamespace
import rbm_b200
from rbm_b200 import lib as L
DEV = "cuda"
V, Ln, d, B = 50, 16, 32, 24
a = SimpleNamespace(model_code="bert", num_items=V, max_len=Ln, device=DEV, model_init_seed=3, bert_num_blocks=2, bert_num_heads=2,
                    bert_hidden_units=d, bert_dropout=0.2, bert_hidden_dropout=0.2, bert_mask_prob=0.2)
common = dict(optimizer="Adam", lr=2e-3, weight_decay=0, momentum=None, decay_step=50, gamma=1.0, num_epochs=1, metric_ks=[10],
              best_metric="NDCG@10", train_batch_size=B, resume_path=None)
m = rbm_b200.model_factory(a)
t = rbm_b200.trainer_factory(SimpleNamespace(**vars(a), **common), m, None, None, None, None)
m.train()
rs = np.random.RandomState(1)
tok = rs.randint(1, V + 1, size=(B, Ln)); lab = np.where(rs.rand(B, Ln) < 0.3, tok, 0)
batch = (torch.from_numpy(np.where(lab != 0, V + 1, tok)).to(DEV), torch.from_numpy(lab).to(DEV))
stage = sys.argv[1]
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        t.optimizer.zero_grad(); loss = t.calculate_loss(batch); loss.backward(); t.optimizer.step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
t.optimizer.zero_grad(set_to_none=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss = t.calculate_loss(batch)
    if stage in ("bwd", "opt"):
        loss.backward()
    if stage == "opt":
        t.optimizer.step()
print("captured", stage)
g.replay(); torch.cuda.synchronize()
print("replayed", stage, loss.item())
