"""Bring-up helpers for the d = 128 / 256 model shapes: `parity` (hidden states and loss against the oracle) and `prof`
(per-entry-point device time of one BERT4Rec step at the BASELINE configs[3] shape on one GPU)."""
import sys, os
from types import SimpleNamespace
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbm_b200
from rbm_b200 import lib as L
from oracle import bert4rec as ob
DEV = "cuda"
def bert_args(V, Ln, d, nb, h, p=0.0, seed=0):
    return SimpleNamespace(model_code="bert", num_items=V, max_len=Ln, device=DEV, model_init_seed=seed, bert_num_blocks=nb,
                           bert_num_heads=h, bert_hidden_units=d, bert_dropout=p, bert_hidden_dropout=p)
if sys.argv[1] == "parity":
    for (V, Ln, d, nb, h, B) in [(900, 200, 256, 1, 4, 3), (900, 200, 256, 4, 4, 3), (900, 64, 256, 1, 4, 3), (900, 200, 128, 1, 2, 3)]:
        rng = np.random.RandomState(V + d)
        model = rbm_b200.model_factory(bert_args(V, Ln, d, nb, h, seed=2))
        model.load_state_dict(ob.random_state_dict(V, Ln, d, nb, seed=5))
        model.to(DEV).train()
        tok = rng.randint(1, V + 1, size=(B, Ln)).astype(np.int64); tok[0, : Ln // 3] = 0
        lab = np.where((rng.rand(B, Ln) < 0.2) & (tok != 0), tok, 0)
        tok = np.where(lab != 0, V + 1, tok)
        t, l = torch.from_numpy(tok), torch.from_numpy(lab)
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        with torch.no_grad():
            hg = model.hidden_states(t).cpu()
            hr = ob.hidden_states(sd, t, nb, h)
            lg = model.loss(t, l).item(); lr = ob.loss(sd, t, l, nb, h).item()
        print((V, Ln, d, nb, h, B), "hidden max err %.3e (scale %.3e)  loss %.7f vs %.7f rel %.2e" % ((hg - hr).abs().max(), hr.abs().max(), lg, lr, abs(lg - lr) / abs(lr)))
else:
    V4, L4, d4, B4 = 1_000_000, 200, 256, 32
    a4 = SimpleNamespace(model_code="bert", num_items=V4, max_len=L4, device=DEV, model_init_seed=0, bert_num_blocks=4,
                         bert_num_heads=4, bert_hidden_units=d4, bert_dropout=0.1, bert_hidden_dropout=0.1, optimizer="Adam", lr=1e-3,
                         weight_decay=0, momentum=None, decay_step=25, gamma=1.0, num_epochs=1, metric_ks=[10], best_metric="NDCG@10",
                         train_batch_size=B4, resume_path=None)
    with torch.device(DEV):
        m4 = rbm_b200.model_factory(a4)
    t4 = rbm_b200.trainer_factory(a4, m4, None, None, None, None)
    m4.train()
    tok = torch.randint(1, V4 + 1, (B4, L4), device=DEV)
    lab = torch.where(torch.rand(B4, L4, device=DEV) < 0.15, tok, torch.zeros_like(tok))
    b = (torch.where(lab != 0, torch.full_like(tok, V4 + 1), tok), lab)
    t4.train_step(b); torch.cuda.synchronize()
    L.profile = {}
    t4.train_step(b)
    prof = L.profile_collect(); L.profile = None
    tot = {k: sum(ms for ms, _ in v) for k, v in prof.items()}
    s = sum(tot.values())
    print("sum of entry points: %.3f ms/step (B=%d)" % (s, B4))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
        print("%6.1f%%  %10.1f us  x%-3d %s" % (100 * v / s, v * 1e3, len(prof[k]), k))
