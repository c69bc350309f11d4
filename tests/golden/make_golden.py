#!/usr/bin/env python
"""Mint golden vectors from the UNMODIFIED reference, run in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (``/root/reference/NerualNetwork/bert4rec&sas4rec``) has no tests or fixtures of its
own, so these vectors -- outputs of the reference's own ``BERTModel`` /
``SASModel`` / ``recalls_ndcgs_and_mrr_for_ks`` / ``optim.Adam`` on seeded inputs -- are what pins the
CPU oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Nothing here is read at
test time except the .npz files; ``/root/reference`` does not exist on the GPU box.
"""
import argparse
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/NerualNetwork/bert4rec&sas4rec"


def ref_imports():
    sys.argv = sys.argv[:1]
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    import models  # noqa
    from models import model_factory
    from trainers.utils import recalls_ndcgs_and_mrr_for_ks
    return model_factory, recalls_ndcgs_and_mrr_for_ks


def bert_args(V, L, d, nb, h, p=0.0, seed=0):
    return SimpleNamespace(model_code="bert", num_items=V, max_len=L, device="cpu", model_init_seed=seed,
                           bert_num_blocks=nb, bert_num_heads=h, bert_hidden_units=d, bert_dropout=p,
                           bert_hidden_dropout=p)


def sas_args(V, L, d, nb, h, p=0.0):
    return SimpleNamespace(model_code="sas", num_items=V, max_len=L, device="cpu", sas_hidden_units=d,
                           sas_num_blocks=nb, sas_heads=h, sas_dropout=p)


def make_bert_batch(rng, B, L, V, mask_prob=0.3):
    """Wire format of BertTrainDataset.__getitem__ (NN/dataloaders/bert.py:77-110)."""
    toks = np.zeros((B, L), np.int64)
    labs = np.zeros((B, L), np.int64)
    for b in range(B):
        n = rng.randint(max(2, L // 3), L + 1) if b else L  # first row unpadded
        seq = rng.randint(1, V + 1, size=n)
        t, l = [], []
        for s in seq:
            pr = rng.rand()
            if pr < mask_prob:
                pr /= mask_prob
                t.append(V + 1 if pr < 0.8 else (rng.randint(1, V + 1) if pr < 0.9 else s))
                l.append(s)
            else:
                t.append(s)
                l.append(0)
        toks[b, L - n:] = t
        labs[b, L - n:] = l
    if labs.sum() == 0:
        labs[0, -1] = toks[0, -1]
    return toks, labs


def make_sas_batch(rng, B, L, V):
    """Wire format of sample_function (NN/dataloaders/sas.py:70-86), negatives may be item 0."""
    seq = np.zeros((B, L), np.int64)
    pos = np.zeros((B, L), np.int64)
    neg = np.zeros((B, L), np.int64)
    for b in range(B):
        n = rng.randint(3, L + 2) if b else L + 1
        train = rng.randint(1, V + 1, size=n)
        pad = L - n + 1
        seq[b, pad:] = train[:-1]
        pos[b, pad:] = train[1:]
        allowed = np.array(sorted(set(range(0, V + 1)) - set(train.tolist())))
        neg[b, pad:] = allowed[rng.randint(0, len(allowed), size=n - 1)]
    return seq, pos, neg


def sd_numpy(model, prefix="sd."):
    return {prefix + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def grads_numpy(model, prefix="grad."):
    return {prefix + k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}


def gen_bert(model_factory, name, V, L, d, nb, h, B, seed, store_sd=True, adam_steps=5):
    rng = np.random.RandomState(seed)
    args = bert_args(V, L, d, nb, h, seed=seed)
    model = model_factory(args)
    if not store_sd:  # cfg-shaped case: weights regenerated from the seed at test time, not stored
        from oracle import bert4rec as ob
        model.load_state_dict(ob.random_state_dict(V, L, d, nb, seed=seed))
    model.train()
    toks, labs = make_bert_batch(rng, B, L, V)
    t, l = torch.from_numpy(toks), torch.from_numpy(labs)
    out = {"tokens": toks, "labels": labs, "cfg": np.array([V, L, d, nb, h, B, seed], np.int64)}
    if store_sd:
        out.update(sd_numpy(model))
    logits = model(t)
    ce = torch.nn.CrossEntropyLoss(ignore_index=0)
    loss = ce(logits.view(-1, logits.size(-1)), l.view(-1))
    model.zero_grad()
    loss.backward()
    out["loss"] = np.float32(loss.item())
    if store_sd:
        out["logits"] = logits.detach().numpy()
        out.update(grads_numpy(model))
    else:
        out["logits_last"] = logits[:, -1, :].detach().numpy()
        out["grad_norms"] = np.array([p.grad.norm().item() for _, p in model.named_parameters()], np.float32)
    # eval-style candidate scores (NN/trainers/bert.py:43-49): seq ends in [MASK]
    ev = toks.copy()
    ev[:, :-1] = toks[:, 1:]
    ev[:, -1] = V + 1
    cands = np.stack([rng.permutation(np.arange(1, V + 1))[:min(V, 21)] for _ in range(B)]).astype(np.int64)
    model.eval()
    with torch.no_grad():
        sc = model(torch.from_numpy(ev))[:, -1, :]
        out["eval_tokens"], out["candidates"] = ev, cands
        out["cand_scores"] = sc.gather(1, torch.from_numpy(cands)).numpy()
        out["scores_last"] = sc.numpy()
    # a few Adam steps on the same batch (NN/trainers/base.py:110-123, :228)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0)
    losses = []
    for _ in range(adam_steps):
        opt.zero_grad()
        lg = model(t)
        ls = ce(lg.view(-1, lg.size(-1)), l.view(-1))
        losses.append(ls.item())
        ls.backward()
        opt.step()
    out["adam_losses"] = np.array(losses, np.float32)
    if store_sd:
        out.update(sd_numpy(model, "sd_after."))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", out["loss"], "adam", losses)


def gen_sas(model_factory, name, V, L, d, nb, h, B, seed, store_sd=True, adam_steps=5):
    rng = np.random.RandomState(seed)
    torch.manual_seed(seed)  # SAS never seeds itself
    model = model_factory(sas_args(V, L, d, nb, h))
    if not store_sd:
        from oracle import sasrec as osr
        model.load_state_dict(osr.random_state_dict(V, L, d, nb, seed=seed))
    model.train()
    seq, pos, neg = make_sas_batch(rng, B, L, V)
    out = {"seq": seq, "pos": pos, "neg": neg, "cfg": np.array([V, L, d, nb, h, B, seed], np.int64)}
    if store_sd:
        out.update(sd_numpy(model))
    bce = torch.nn.BCEWithLogitsLoss()

    def step_loss():
        pl, nl = model(seq, pos, neg)
        idx = np.where(pos != 0)
        return pl, nl, bce(pl[idx], torch.ones_like(pl)[idx]) + bce(nl[idx], torch.zeros_like(nl)[idx])

    pl, nl, loss = step_loss()
    model.zero_grad()
    loss.backward()
    out["pos_logits"], out["neg_logits"] = pl.detach().numpy(), nl.detach().numpy()
    out["loss"] = np.float32(loss.item())
    if store_sd:
        out.update(grads_numpy(model))
    else:
        out["grad_norms"] = np.array([p.grad.norm().item() for _, p in model.named_parameters()], np.float32)
    cands = np.stack([rng.permutation(np.arange(1, V + 1))[:min(V, 21)] for _ in range(B)]).astype(np.int64)
    model.eval()
    with torch.no_grad():
        out["candidates"] = cands
        out["cand_scores"] = model.predict(seq, cands).numpy()
        allc = np.tile(np.arange(1, V + 1, dtype=np.int64), (B, 1))
        out["scores_full"] = model.predict(seq, allc).numpy()
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0)
    losses = []
    for _ in range(adam_steps):
        opt.zero_grad()
        _, _, ls = step_loss()
        losses.append(ls.item())
        ls.backward()
        opt.step()
    out["adam_losses"] = np.array(losses, np.float32)
    if store_sd:
        out.update(sd_numpy(model, "sd_after."))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", out["loss"], "adam", losses)


def gen_metrics(fn, name="metrics"):
    rng = np.random.RandomState(7)
    out = {}
    for tag, (B, C) in {"c101": (64, 101), "c3416": (16, 3416)}.items():
        scores = rng.randn(B, C).astype(np.float32)
        labels = np.zeros((B, C), np.int64)
        labels[:, 0] = 1
        # make a decent share of positives land in the top-20
        boost = rng.rand(B) < 0.6
        scores[boost, 0] += 2.5
        ks = [1, 5, 10, 20]
        m = fn(torch.from_numpy(scores), torch.from_numpy(labels), ks)
        out[tag + ".scores"], out[tag + ".labels"] = scores, labels
        out[tag + ".keys"] = np.array(sorted(m.keys()))
        out[tag + ".vals"] = np.array([m[k] for k in sorted(m.keys())], np.float64)
        rank = (-torch.from_numpy(scores)).argsort(dim=1)[:, :20].numpy()
        out[tag + ".rank20"] = rank
    # multi-positive case (labels with 1..3 positives)
    B, C = 32, 50
    scores = rng.randn(B, C).astype(np.float32)
    labels = (rng.rand(B, C) < 0.05).astype(np.int64)
    labels[:, 0] = 1
    m = fn(torch.from_numpy(scores), torch.from_numpy(labels), [1, 5, 10])
    out["multi.scores"], out["multi.labels"] = scores, labels
    out["multi.keys"] = np.array(sorted(m.keys()))
    out["multi.vals"] = np.array([m[k] for k in sorted(m.keys())], np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written")


def gen_scatter_adam(name="scatter_adam"):
    """embedding_dense_backward and optim.Adam of the installed torch (the reference's L0 substrate)."""
    rng = np.random.RandomState(11)
    out = {}
    V, d, n = 53, 16, 400
    idx = rng.zipf(1.3, size=n) % V
    rows = rng.randn(n, d).astype(np.float32)
    emb = torch.nn.Embedding(V, d, padding_idx=0)
    y = emb(torch.from_numpy(idx.astype(np.int64)))
    y.backward(torch.from_numpy(rows))
    out["scatter.idx"], out["scatter.rows"], out["scatter.grad"] = idx.astype(np.int64), rows, emb.weight.grad.numpy().copy()
    p = torch.nn.Parameter(torch.from_numpy(rng.randn(300).astype(np.float32)))
    out["adam.p0"] = p.detach().numpy().copy()
    opt = torch.optim.Adam([p], lr=1e-3)
    gs = rng.randn(4, 300).astype(np.float32)
    gs[2, :100] = 0  # rows with zero grad still move (dense Adam semantics)
    ps = []
    for i in range(4):
        p.grad = torch.from_numpy(gs[i].copy())
        opt.step()
        ps.append(p.detach().numpy().copy())
    out["adam.grads"], out["adam.ps"] = gs, np.stack(ps)
    st = opt.state[p]
    out["adam.m"], out["adam.v"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written")


def gen_batches(name="batches"):
    """Training-batch construction of the REFERENCE dataloaders with the contract's random numbers injected:
    ``BertTrainDataset.__getitem__`` (NN/dataloaders/bert.py:77-110) gets a fake ``rng`` whose ``rand`` / ``randint`` return
    the Philox words of (batch row, output column); ``sample_function`` (NN/dataloaders/sas.py:73-90) runs with
    ``np.random.randint`` patched the same way and a queue that stops it after one batch."""
    from oracle.batches import rbm_philox
    import dataloaders.bert as ref_bert
    import dataloaders.sas as ref_sas

    rs = np.random.RandomState(11)
    V, U = 53, 12
    hist = [list(rs.randint(1, V + 1, size=n)) for n in [0, 1, 2, 3, 7, 8, 9, 13, 20, 40, 5, 60]]
    hist[10] = [4, 4, 4, 9, 4]  # repeats inside a window
    out = {"hist_ptr": np.cumsum([0] + [len(h) for h in hist]).astype(np.int64),
           "hist_items": np.array([i for h in hist for i in h], np.int64), "num_items": np.int64(V)}
    seed, site = 1234567, (1 << 62) + 5

    # ---- BERT Cloze
    for L, mask_prob, tag in [(8, 0.4, "bert_L8"), (16, 1.0, "bert_L16_all"), (8, 0.0, "bert_L8_none")]:
        users = np.array([3, 4, 5, 6, 7, 8, 9, 1, 0, 11, 10, 4], np.int64)

        class FakeRng:
            def __init__(self, b, n):
                self.b, self.n, self.k = b, n, -1

            def _words(self):
                p = L - self.n + self.k  # output column of the element being drawn for (negative: truncated away)
                if p < 0:
                    return 0xFFFFFFFF, 0
                r = rbm_philox(seed, site, self.b * 128 + (p >> 1))
                return r[2 * (p & 1)], r[2 * (p & 1) + 1]

            def rand(self):
                self.k += 1
                return self._words()[0] / 4294967296.0

            def randint(self, lo, hi):
                return lo + ((self._words()[1] * (hi - lo)) >> 32)

        toks, labs = [], []
        for b, u in enumerate(users):
            ds = ref_bert.BertTrainDataset({0: hist[u]}, L, mask_prob, V + 1, V, FakeRng(b, len(hist[u])))
            t, l = ds[0]
            toks.append(t.numpy()); labs.append(l.numpy())
        out[tag + ".users"], out[tag + ".mask_prob"] = users, np.float64(mask_prob)
        out[tag + ".tokens"], out[tag + ".labels"] = np.stack(toks), np.stack(labs)

    # ---- SASRec (seq, pos, neg)
    for L, tag in [(8, "sas_L8"), (50, "sas_L50")]:
        users = np.array([2, 3, 4, 5, 6, 7, 8, 9, 11, 10], np.int64)  # histories of >= 2 items (the reference indexes train[-1])
        state = {"b": -1, "calls": 0}
        real_randint = np.random.randint

        def fake_randint(lo, hi=None, size=None):
            if size is None:  # the user draw of sample(): users in the given order
                state["b"] += 1
                return state["b"] if state["b"] < len(users) else 0
            b, n = state["b"], min(len(hist[users[state["b"]]]), L)
            pad = L - n + 1
            vals = []
            for k in range(size):
                p = pad + k
                vals.append(lo + ((rbm_philox(seed, site, b * 64 + (p >> 2))[p & 3] * (hi - lo)) >> 32))
            return np.array(vals, np.int64)

        class OneBatch(Exception):
            pass

        class Q:
            def put(self, z):
                self.z = [list(x) for x in z]
                raise OneBatch()

        q = Q()
        ref_sas.np.random.randint = fake_randint
        try:
            ref_sas.sample_function([hist[u] for u in users], V, len(users), L, q)
        except OneBatch:
            pass
        finally:
            ref_sas.np.random.randint = real_randint
        out[tag + ".users"] = users
        out[tag + ".seq"], out[tag + ".pos"], out[tag + ".neg"] = (np.array(x, np.int64) for x in q.z)
    out["seed"], out["site"] = np.uint64(seed), np.uint64(site)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name)


def gen_eval_batches(name="eval_batches"):
    """Evaluation negatives and evaluation batches of the REFERENCE classes with the contract's random numbers injected:
    ``RandomNegativeSampler`` / ``PopularNegativeSampler.generate_negative_samples`` (NN/dataloaders/negative_samplers/) run
    with ``np.random.choice`` replaced by draws from the Philox stream of (user, attempt) -- for the popularity sampler the
    fake returns ``sample_size`` DISTINCT successive inverse-CDF draws, which is what ``choice(replace=False, p=...)`` means --
    and ``BertEvalDataset`` / ``SASEvalDataset.__getitem__`` (NN/dataloaders/bert.py:116-142, sas.py:124-153) run as they are."""
    from oracle.batches import rbm_philox, NEG_MAX_ATTEMPTS
    import dataloaders.negative_samplers.random as ref_rand
    import dataloaders.negative_samplers.popular as ref_pop
    import dataloaders.bert as ref_bert
    import dataloaders.sas as ref_sas

    rs = np.random.RandomState(5)
    V, U, S = 61, 9, 12
    zipf = 1.0 / np.arange(1, V + 1)
    zipf /= zipf.sum()
    lens = [3, 4, 6, 9, 14, 22, 30, 5, 40]
    train = {u: [int(i) for i in rs.choice(V, size=n, p=zipf) + 1] for u, n in enumerate(lens)}
    val = {u: [int(rs.randint(1, V + 1))] for u in range(U)}
    test = {u: [int(rs.randint(1, V + 1))] for u in range(U)}
    seed, site = 424242, 1 << 41
    counts = np.zeros(V + 1, np.int64)
    for d in (train, val, test):
        for u in range(U):
            np.add.at(counts, np.array(d[u]), 1)
    cdf = [int(c) for c in np.cumsum(counts[1:])]
    state = {"u": 0, "t": 0}

    def fake_trange(a, b):
        for u in range(a, b):
            state["u"], state["t"] = u, 0
            yield u

    def r64():
        r = rbm_philox(seed, site, state["u"] * NEG_MAX_ATTEMPTS + state["t"])
        state["t"] += 1
        return (r[0] << 32) | r[1]

    def fake_choice(a, size=None, replace=True, p=None):
        if size is None:  # random.py:29,31
            return (r64() * int(a)) >> 64
        got = []  # popular.py:33
        while len(got) < size:
            x = (r64() * cdf[-1]) >> 64
            item = 1 + next(i for i in range(V) if cdf[i] > x)
            if item not in got:
                got.append(item)
        return np.array(got)

    out = {"num_items": np.int64(V), "sample_size": np.int64(S), "seed": np.uint64(seed), "site": np.uint64(site),
           "pop_counts": counts[1:].copy()}
    for k, d in (("train", train), ("val", val), ("test", test)):
        out[k + "_ptr"] = np.cumsum([0] + [len(d[u]) for u in range(U)]).astype(np.int64)
        out[k + "_items"] = np.array([i for u in range(U) for i in d[u]], np.int64)
    real_choice = np.random.choice
    negs = {}
    for tag, mod, cls in (("random", ref_rand, ref_rand.RandomNegativeSampler), ("popular", ref_pop, ref_pop.PopularNegativeSampler)):
        real_trange = mod.trange
        mod.trange = fake_trange
        mod.np.random.choice = fake_choice
        try:
            ns = cls(train, val, test, U, V, S, 7, "/tmp", "golden").generate_negative_samples()
        finally:
            mod.trange = real_trange
            mod.np.random.choice = real_choice
        negs[tag] = np.array([[int(i) for i in ns[u]] for u in range(U)], np.int64)
        out["neg_" + tag] = negs[tag]
    # evaluation batches (validation protocol: u2seq = train, answers = val)
    neg_dict = {u: [int(i) for i in negs["random"][u]] for u in range(U)}
    for L in (8, 16):
        bd = ref_bert.BertEvalDataset(train, val, L, V + 1, neg_dict)
        sd = ref_sas.SASEvalDataset(train, val, L, neg_dict)
        rows_b, rows_s = [bd[u] for u in range(U)], [sd[u] for u in range(U)]
        for j, nm in enumerate(("seq", "cand", "labels")):
            out["bert_L%d.%s" % (L, nm)] = np.stack([np.asarray(r[j]) for r in rows_b]).astype(np.int64)
            out["sas_L%d.%s" % (L, nm)] = np.stack([np.asarray(r[j]) for r in rows_s]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name)


def gen_partition(name="partition"):
    """``data_partition`` of the REFERENCE (NN/dataloaders/__init__.py:17-66) on a small interactions file; the function opens
    ``Data/<fname>`` next to the (read-only) reference tree, so its ``open`` is redirected to an in-memory file."""
    import io
    import dataloaders as ref_dl
    rs = np.random.RandomState(21)
    lens = {7: 1, 3: 2, 12: 3, 5: 9, 40: 50, 41: 51, 2: 64, 9: 137, 30: 8, 31: 230}
    lines = []
    pending = {u: list(rs.randint(1, 400, size=n)) for u, n in lens.items()}
    users = list(lens)
    while any(pending.values()):  # interleave the users' lines: grouping must not depend on contiguity
        u = users[rs.randint(len(users))]
        if pending[u]:
            lines.append("%d %d" % (u, pending[u].pop(0)))
    text = "\n".join(lines) + "\n"
    out = {"text": np.frombuffer(text.encode(), dtype=np.uint8)}
    real_open = getattr(ref_dl, "open", None)
    ref_dl.open = lambda path, mode="r": io.StringIO(text)
    try:
        for L, prop in ((8, 0.3), (50, 0.3), (50, -1.0), (200, 0.3)):
            tr, va, te, n, V = ref_dl.data_partition("whatever.txt", L, prop)
            tag = "L%d_p%s" % (L, str(prop).replace(".", "").replace("-", "m"))
            out[tag + ".train_ptr"] = np.cumsum([0] + [len(r) for r in tr]).astype(np.int64)
            out[tag + ".train"] = np.array([i for r in tr for i in r], np.int64)
            out[tag + ".valid"] = np.array([r[0] for r in va], np.int64)
            out[tag + ".test"] = np.array([r[0] for r in te], np.int64)
            out[tag + ".n"], out[tag + ".V"] = np.int64(n), np.int64(V)
    finally:
        if real_open is None:
            del ref_dl.open
        else:
            ref_dl.open = real_open
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, help="regenerate one fixture family only (eval_batches)")
    args = ap.parse_args()
    torch.set_num_threads(1)  # fixed reduction order in the generated vectors
    model_factory, metric_fn = ref_imports()
    if args.only == "eval_batches":
        return gen_eval_batches()
    if args.only == "partition":
        return gen_partition()
    if args.only == "wide":  # the d = 128 / 256 model shapes of BASELINE configs[2] / configs[3] (small catalogues)
        gen_bert(model_factory, "bert_cfg4s", V=900, L=200, d=256, nb=4, h=4, B=3, seed=2, store_sd=False, adam_steps=1)
        return gen_sas(model_factory, "sas_cfg3s", V=1200, L=50, d=128, nb=2, h=2, B=8, seed=2, store_sd=False, adam_steps=1)
    gen_bert(model_factory, "bert_tiny", V=37, L=8, d=16, nb=2, h=2, B=4, seed=0)
    gen_bert(model_factory, "bert_odd", V=101, L=13, d=32, nb=1, h=4, B=3, seed=3)
    gen_bert(model_factory, "bert_cfg2", V=3416, L=200, d=64, nb=2, h=2, B=4, seed=1, store_sd=False, adam_steps=2)
    gen_sas(model_factory, "sas_tiny", V=37, L=8, d=16, nb=2, h=2, B=4, seed=0)
    gen_sas(model_factory, "sas_odd", V=101, L=13, d=32, nb=1, h=1, B=3, seed=3)
    gen_sas(model_factory, "sas_cfg1", V=3416, L=50, d=64, nb=2, h=1, B=8, seed=1, store_sd=False, adam_steps=2)
    gen_metrics(metric_fn)
    gen_scatter_adam()
    gen_batches()
    gen_eval_batches()
    gen_partition()
    gen_bert(model_factory, "bert_cfg4s", V=900, L=200, d=256, nb=4, h=4, B=3, seed=2, store_sd=False, adam_steps=1)
    gen_sas(model_factory, "sas_cfg3s", V=1200, L=50, d=128, nb=2, h=2, B=8, seed=2, store_sd=False, adam_steps=1)


if __name__ == "__main__":
    main()
