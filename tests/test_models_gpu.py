"""Model-level parity against golden vectors minted from the reference (tests/golden/make_golden.py) and the oracle."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import rbm_b200
from rbm_b200 import ops
from oracle import bert4rec as ob, sasrec as osr, metrics as om

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def load(name):
    z = np.load(os.path.join(G, name + ".npz"))
    return {k: z[k] for k in z.files}


def bert_args(V, Ln, d, nb, h, p=0.0, seed=0):
    return SimpleNamespace(model_code="bert", num_items=V, max_len=Ln, device=DEV, model_init_seed=seed, bert_num_blocks=nb,
                           bert_num_heads=h, bert_hidden_units=d, bert_dropout=p, bert_hidden_dropout=p)


def sas_args(V, Ln, d, nb, h, p=0.0):
    return SimpleNamespace(model_code="sas", num_items=V, max_len=Ln, device=DEV, sas_hidden_units=d, sas_num_blocks=nb,
                           sas_heads=h, sas_dropout=p)


def sd_of(z, prefix="sd."):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in z.items() if k.startswith(prefix)}


def relclose(a, b, rel=1e-3, floor=1e-4, msg=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), floor)
    assert np.abs(a - b).max() <= rel * scale, "%s: max err %g vs scale %g" % (msg, np.abs(a - b).max(), scale)


@pytest.mark.parametrize("label_rows", [False, True])
@pytest.mark.parametrize("name", ["bert_tiny", "bert_odd"])
def test_bert_vs_reference_golden(name, label_rows):
    """``label_rows``: the final block's output projection / LayerNorm / feed-forward on the labelled rows only (forced here
    whatever the label share; off = every row) -- same goldens, same tolerances."""
    z = load(name)
    V, Ln, d, nb, h, B, seed = z["cfg"].tolist()
    model = rbm_b200.model_factory(bert_args(V, Ln, d, nb, h, seed=seed))
    model.load_state_dict(sd_of(z))
    model.to(DEV).train()
    model.LABEL_ROWS_MAX_FRACTION = 1.0 if label_rows else -1.0
    t, l = torch.from_numpy(z["tokens"]), torch.from_numpy(z["labels"])
    logits = model(t.to(DEV))
    assert logits.shape == (B, Ln, V + 1) and logits.is_contiguous()
    relclose(logits.detach().cpu().numpy(), z["logits"], 1e-4, msg="logits")
    loss = model.loss(t, l)
    assert abs(loss.item() - float(z["loss"])) < 1e-5 * abs(float(z["loss"])) + 1e-6
    loss.backward()
    for k, p in model.named_parameters():
        relclose(p.grad.cpu().numpy(), z["grad." + k], 1e-3, msg=k)
    # the compatibility forward + torch CE gives the same loss and gradients through the same kernels
    model.zero_grad()
    lg = model(t.to(DEV))
    l2 = torch.nn.functional.cross_entropy(lg.view(-1, V + 1), l.to(DEV).view(-1), ignore_index=0)
    l2.backward()
    assert abs(l2.item() - float(z["loss"])) < 1e-5 * abs(float(z["loss"])) + 1e-6
    relclose(model.out.weight.grad.cpu().numpy(), z["grad.out.weight"], 1e-3, msg="out.weight via forward()")
    # evaluation scoring
    model.eval()
    with torch.no_grad():
        cs = model.candidate_scores(torch.from_numpy(z["eval_tokens"]), torch.from_numpy(z["candidates"]))
        relclose(cs.cpu().numpy(), z["cand_scores"], 1e-4, msg="cand_scores")
        vals, ids = model.full_catalogue_topk(torch.from_numpy(z["eval_tokens"]), 10)
        ref_v, ref_i = om.topk_canonical(z["scores_last"][:, 1:], 10, id_offset=1)
        relclose(vals.cpu().numpy(), ref_v, 1e-4, msg="topk scores")
        np.testing.assert_array_equal(ids.cpu().numpy(), ref_i)


def test_bert_adam_trajectory():
    z = load("bert_tiny")
    V, Ln, d, nb, h, B, seed = z["cfg"].tolist()
    args = bert_args(V, Ln, d, nb, h, seed=seed)
    args.update = None
    model = rbm_b200.model_factory(args)
    model.load_state_dict(sd_of(z))
    targs = SimpleNamespace(**vars(args), optimizer="Adam", lr=1e-3, weight_decay=0, momentum=None, decay_step=50, gamma=0.95,
                            num_epochs=1, metric_ks=[1, 5, 10], best_metric="NDCG@10", train_batch_size=B, resume_path=None)
    trainer = rbm_b200.trainer_factory(targs, model, None, None, None, None)
    model.train()
    batch = (torch.from_numpy(z["tokens"]), torch.from_numpy(z["labels"]))
    losses = [trainer.train_step(batch).item() for _ in range(len(z["adam_losses"]))]
    np.testing.assert_allclose(losses, z["adam_losses"], rtol=1e-4)
    after = sd_of(z, "sd_after.")
    for k, v in model.state_dict().items():
        if k.endswith("linear_layers.1.bias"):
            continue  # analytically-zero gradient amplified by Adam (see tests/test_oracle_golden.py)
        relclose(v.cpu().numpy(), after[k].numpy(), 2e-3, msg=k)


@pytest.mark.parametrize("live_rows", [False, True])
@pytest.mark.parametrize("name", ["sas_tiny", "sas_odd"])
def test_sas_vs_reference_golden(name, live_rows):
    """``live_rows``: the token-wise layers run on the non-padding rows only (forced here whatever the padding share; off = the
    dense path) -- same goldens, same tolerances, logits at the padding positions included."""
    z = load(name)
    V, Ln, d, nb, h, B, seed = z["cfg"].tolist()
    model = rbm_b200.model_factory(sas_args(V, Ln, d, nb, h))
    model.load_state_dict(sd_of(z))
    model.to(DEV).train()
    model.LIVE_ROWS_MAX_FRACTION = 1.0 if live_rows else -1.0
    pl, nl = model(z["seq"], z["pos"], z["neg"])  # numpy in, like the reference
    relclose(pl.detach().cpu().numpy(), z["pos_logits"], 1e-4, msg="pos_logits")
    relclose(nl.detach().cpu().numpy(), z["neg_logits"], 1e-4, msg="neg_logits")
    loss = model.loss(z["seq"], z["pos"], z["neg"])
    assert abs(loss.item() - float(z["loss"])) < 1e-5 * abs(float(z["loss"])) + 1e-6
    loss.backward()
    for k, p in model.named_parameters():
        relclose(p.grad.cpu().numpy(), z["grad." + k], 1e-3, msg=k)
    model.eval()
    with torch.no_grad():
        relclose(model.predict(z["seq"], z["candidates"]).cpu().numpy(), z["cand_scores"], 1e-4, msg="predict")
        vals, ids = model.full_catalogue_topk(z["seq"], 10)
        ref_v, ref_i = om.topk_canonical(z["scores_full"], 10, id_offset=1)
        relclose(vals.cpu().numpy(), ref_v, 1e-4, msg="topk scores")
        np.testing.assert_array_equal(ids.cpu().numpy(), ref_i)


def test_sas_adam_trajectory():
    z = load("sas_tiny")
    V, Ln, d, nb, h, B, seed = z["cfg"].tolist()
    args = sas_args(V, Ln, d, nb, h)
    model = rbm_b200.model_factory(args)
    model.load_state_dict(sd_of(z))
    targs = SimpleNamespace(**vars(args), optimizer="Adam", lr=1e-3, weight_decay=0, momentum=None, decay_step=50, gamma=0.95,
                            num_epochs=1, metric_ks=[1, 5, 10], best_metric="NDCG@10", train_batch_size=B, resume_path=None, l2_emb=0.0)
    trainer = rbm_b200.trainer_factory(targs, model, None, None, None, None)
    model.train()
    losses = [trainer.train_step((z["seq"], z["pos"], z["neg"])).item() for _ in range(len(z["adam_losses"]))]
    np.testing.assert_allclose(losses, z["adam_losses"], rtol=1e-4)


def test_cfg_shaped_losses():
    """cfg2-shaped BERT4Rec and cfg1-shaped SASRec batches: loss / grad norms / scores vs the reference goldens."""
    z = load("bert_cfg2")
    V, Ln, d, nb, h, B, seed = z["cfg"].tolist()
    model = rbm_b200.model_factory(bert_args(V, Ln, d, nb, h, seed=seed))
    model.load_state_dict(ob.random_state_dict(V, Ln, d, nb, seed=seed))
    model.to(DEV).train()
    loss = model.loss(torch.from_numpy(z["tokens"]), torch.from_numpy(z["labels"]))
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    gn = np.array([p.grad.norm().item() for _, p in model.named_parameters()])
    np.testing.assert_allclose(gn, z["grad_norms"], rtol=2e-3, atol=1e-7)
    model.eval()
    with torch.no_grad():
        vals, ids = model.full_catalogue_topk(torch.from_numpy(z["eval_tokens"]), 10)
    ref_v, ref_i = om.topk_canonical(z["scores_last"][:, 1:], 10, id_offset=1)
    relclose(vals.cpu().numpy(), ref_v, 1e-4)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_i)

    z = load("sas_cfg1")
    V, Ln, d, nb, h, B, seed = z["cfg"].tolist()
    model = rbm_b200.model_factory(sas_args(V, Ln, d, nb, h))
    model.load_state_dict(osr.random_state_dict(V, Ln, d, nb, seed=seed))
    model.to(DEV).train()
    loss = model.loss(z["seq"], z["pos"], z["neg"])
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    gn = np.array([p.grad.norm().item() for _, p in model.named_parameters()])
    np.testing.assert_allclose(gn, z["grad_norms"], rtol=2e-3, atol=1e-7)
    model.eval()
    with torch.no_grad():
        vals, ids = model.full_catalogue_topk(z["seq"], 10)
    ref_v, ref_i = om.topk_canonical(z["scores_full"], 10, id_offset=1)
    relclose(vals.cpu().numpy(), ref_v, 1e-4)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_i)


@pytest.mark.parametrize("label_rows,Ln,d", [(True, 8, 16), (False, 8, 16), (True, 40, 64)])
def test_training_mode_dropout_against_injected_masks(label_rows, Ln, d):
    """Training-mode BERT4Rec: the oracle consumes the Philox masks the kernels used (exported per site).  ``label_rows``: the
    final block's output projection / LayerNorm / feed-forward on the labelled rows only (the default when at most half of the
    positions are labelled) or on every row; at d_k = 32 (d = 64) the final block's attention also takes the labelled positions'
    queries only, compacted per sequence."""
    from oracle.common import DropoutPlan
    if label_rows and (os.environ.get("RBM_BERT_LABEL_ROWS", "1") == "0" or (d == 64 and os.environ.get("RBM_BERT_LABEL_QUERIES", "1") == "0")):
        pytest.skip("the labelled-row path is switched off by the environment")
    V, nb, h, B, p = 37, 2, 2, 4, 0.25
    model = rbm_b200.model_factory(bert_args(V, Ln, d, nb, h, p=p, seed=1)).to(DEV).train()
    model.dropout_seed = 4242
    model.LABEL_ROWS_MAX_FRACTION = 1.0 if label_rows else -1.0
    rng = np.random.RandomState(0)
    tok = torch.from_numpy(rng.randint(1, V + 2, size=(B, Ln)).astype(np.int64))
    tok[0, :3] = 0
    lab = torch.from_numpy(np.where(rng.rand(B, Ln) < (0.4 if Ln == 8 else 0.25), rng.randint(1, V + 1, size=(B, Ln)), 0).astype(np.int64))
    step0 = model._step
    loss = model.loss(tok, lab)
    loss.backward()
    base = step0 * 64
    masks = {0: ops.dropout_mask(B * Ln * d, p, 4242, base, DEV).cpu()}
    labelled = np.flatnonzero(lab.numpy().reshape(-1))
    compact = len(labelled) <= model.LABEL_ROWS_MAX_FRACTION * B * Ln  # the final block then runs on the labelled rows only
    per_seq = (lab != 0).sum(1)
    cap, lq = model._capacities(len(labelled), int(per_seq.max()))
    compact_q = compact and d // h == 32 and lq < Ln  # the final block's attention on per-sequence compacted queries
    assert compact_q == (label_rows and d == 64)
    for b in range(nb):
        s = ob.block_sites(b)
        am = ops.dropout_mask_attn(B * h * Ln, Ln, p, 4242, base + s["attn"], DEV).cpu()
        if compact_q and b == nb - 1:
            # Philox stream indexed by (sequence-head, ordinal of the labelled position within its sequence, key)
            am, src = torch.ones_like(am).view(B, h, Ln, Ln), am.view(B, h, Ln, Ln)
            for bb in range(B):
                for o, pos in enumerate(np.flatnonzero(lab[bb].numpy())):
                    am[bb, :, pos, :] = src[bb, :, o, :]
        masks[s["attn"]] = am
        for key, w in (("sub_in", d), ("ffn", 4 * d), ("sub_out", d), ("block", d)):
            if compact and b == nb - 1:
                # element-wise sites of the final block: Philox stream indexed by (labelled-row ordinal, column); the other rows of
                # that block never reach the loss, whatever mask the oracle applies to them
                mc = ops.dropout_mask(cap * w, p, 4242, base + s[key], DEV).cpu().reshape(cap, w)
                full = torch.ones(B * Ln, w, dtype=mc.dtype)
                full[torch.from_numpy(labelled)] = mc[: len(labelled)]
                masks[s[key]] = full
            else:
                masks[s[key]] = ops.dropout_mask(B * Ln * w, p, 4242, base + s[key], DEV).cpu()
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ref = ob.loss(sd, tok, lab, nb, h, p_attn=p, p_hidden=p, drop=DropoutPlan(True, masks))
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    for k, prm in model.named_parameters():
        gr = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
        relclose(prm.grad.cpu().numpy(), gr.numpy(), 1e-3, msg=k)
    # a second forward draws a different mask stream
    assert model.loss(tok, lab).item() != loss.item()


def test_training_from_device_side_loaders():
    """Both models train from batches built on the device (rbm_bert_cloze_batch / rbm_sas_train_batch): the trainers consume
    the loaders' output unchanged and the loss goes down on a learnable toy stream (every user walks the item ring)."""
    from rbm_b200.dataloaders import DeviceBertTrainLoader, DeviceSasTrainLoader
    V, U, Ln = 40, 256, 16
    rs = np.random.RandomState(0)
    hist = []
    for u in range(U):
        s = rs.randint(1, V + 1)
        hist.append([(s + k - 1) % V + 1 for k in range(rs.randint(6, 30))])  # next item = current + 1 (mod V)
    common = dict(optimizer="Adam", lr=5e-3, weight_decay=0, momentum=None, decay_step=50, gamma=1.0, num_epochs=1, metric_ks=[10],
                  best_metric="NDCG@10", train_batch_size=64, resume_path=None)
    # BERT4Rec
    a = bert_args(V, Ln, 32, 1, 2, p=0.0, seed=1)
    model = rbm_b200.model_factory(a)
    tr = rbm_b200.trainer_factory(SimpleNamespace(**vars(a), **common), model, None, None, None, None)
    model.train()
    loader = DeviceBertTrainLoader(hist, Ln, 0.2, V, 64, DEV, seed=3)
    assert len(loader) == 4
    losses = [tr.train_step(b).item() for _ in range(12) for b in loader]
    assert np.isfinite(losses).all() and np.mean(losses[-8:]) < 0.7 * np.mean(losses[:8]), (losses[:8], losses[-8:])
    # SASRec
    a = sas_args(V, Ln, 32, 1, 1, p=0.0)
    model = rbm_b200.model_factory(a)
    tr = rbm_b200.trainer_factory(SimpleNamespace(**vars(a), l2_emb=0.0, **common), model, None, None, None, None)
    model.train()
    loader = DeviceSasTrainLoader(hist, Ln, V, 64, DEV, seed=3)
    assert len(loader) == 4
    losses = [tr.train_step(b).item() for _ in range(12) for b in loader]
    assert np.isfinite(losses).all() and np.mean(losses[-8:]) < 0.8 * np.mean(losses[:8]), (losses[:8], losses[-8:])


@pytest.mark.parametrize("kind", ["bert", "sas"])
def test_cuda_graph_train_step_equals_eager(kind):
    """The captured train step (one CUDA graph, device-side step counter for dropout sites and Adam's bias correction)
    reproduces the eager steps bit for bit: same losses, same parameters, dropout on."""
    V, Ln, d, B = 50, 16, 32, 24
    rs = np.random.RandomState(1)
    if kind == "bert":
        mk = lambda: rbm_b200.model_factory(bert_args(V, Ln, d, 2, 2, p=0.2, seed=3))
        def batch(i):
            t = rs.randint(1, V + 1, size=(B, Ln)); t[:, : i % 5] = 0
            l = np.where(rs.rand(B, Ln) < 0.3, t, 0)
            return torch.from_numpy(np.where(l != 0, V + 1, t)).to(DEV), torch.from_numpy(l).to(DEV)
        extra = {}
    else:
        mk = lambda: rbm_b200.model_factory(sas_args(V, Ln, d, 2, 1, p=0.2))
        def batch(i):
            s = rs.randint(1, V + 1, size=(B, Ln)); s[:, : 1 + i % 5] = 0
            p_ = np.where(s != 0, rs.randint(1, V + 1, size=(B, Ln)), 0)
            n_ = np.where(s != 0, rs.randint(1, V + 1, size=(B, Ln)), 0)
            return tuple(torch.from_numpy(x).to(DEV) for x in (s, p_, n_))
        extra = {"l2_emb": 0.0}
    common = dict(optimizer="Adam", lr=2e-3, weight_decay=0, momentum=None, decay_step=50, gamma=1.0, num_epochs=1, metric_ks=[10],
                  best_metric="NDCG@10", train_batch_size=B, resume_path=None, **extra)
    batches = [batch(i) for i in range(6)]
    torch.manual_seed(0)
    m1 = mk()
    m2 = mk()
    m2.load_state_dict(m1.state_dict())
    m2.dropout_seed = m1.dropout_seed  # (by default the process seed at construction time)
    a = bert_args(V, Ln, d, 2, 2, p=0.2, seed=3) if kind == "bert" else sas_args(V, Ln, d, 2, 1, p=0.2)
    t1 = rbm_b200.trainer_factory(SimpleNamespace(**vars(a), **common), m1, None, None, None, None)
    t2 = rbm_b200.trainer_factory(SimpleNamespace(**vars(a), **common), m2, None, None, None, None)
    m1.train(); m2.train()
    for m in (m1, m2):  # the row-compacting paths size their tensors per batch (eager) / per capture (graph): bit-identity across
        m.LIVE_ROWS_MAX_FRACTION = m.LABEL_ROWS_MAX_FRACTION = -1.0  # capacities is test_live_rows_cuda_graph's subject
    eager = [t1.train_step(b).item() for b in batches]
    t2.capture_train_step(batches[0])
    graphed = [t2.train_step(b).item() for b in batches]
    t2.release_train_graph()
    assert eager == graphed, (eager, graphed)
    for (k, v1), (_, v2) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(v1, v2), k
    # and eager steps continue seamlessly after the graph is released
    nb = batch(7)
    assert t1.train_step(nb).item() == t2.train_step(nb).item()
    # a batch of another shape while the graph is active (the ragged last batch of an epoch: the loaders, like the reference's
    # DataLoader, have no drop_last) takes one eager step at the right position of the dropout / Adam streams -- also the
    # 1-row remainder that copy_ would silently broadcast -- and the checkpoint written in graph mode carries the real steps
    t2.capture_train_step(batches[0])
    seq = [batches[1], tuple(x[:5] for x in batches[2]), batches[3], tuple(x[:1] for x in batches[4]), batches[5]]
    e2 = [t1.train_step(b).item() for b in seq]
    g2 = [t2.train_step(b).item() for b in seq]
    assert e2 == g2, (e2, g2)
    sd1, sd2 = t1._create_state_dict(), t2._create_state_dict()
    assert sd1["dropout_step"] == sd2["dropout_step"] == m1._step
    for k_, st in sd1["optimizer_state_dict"]["state"].items():
        assert float(st["step"]) == float(sd2["optimizer_state_dict"]["state"][k_]["step"])
    t2.release_train_graph()
    for (k, v1), (_, v2) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(v1, v2), k
    assert t1.train_step(nb).item() == t2.train_step(nb).item()


@pytest.mark.parametrize("kind,V,Ln,d,nb,h,B", [("sas", 1200, 50, 128, 2, 2, 16),     # BASELINE configs[2] model shape
                                                 ("bert", 900, 200, 256, 4, 4, 3),     # BASELINE configs[3] model shape (small V)
                                                 ("bert", 300, 64, 128, 1, 2, 5)])
def test_wide_models_vs_oracle(kind, V, Ln, d, nb, h, B):
    """d = 128 / 256 models (column-group tcgen05 Linear, d_k = 64 attention, wide cross-entropy): loss and EVERY gradient
    against the oracle on the same weights and batch (dropout 0); tolerance 1e-3 of each tensor's scale; loss 1e-4 relative
    (measured 1e-5 at d = 256, where K = 1024 contractions and logits of magnitude ~20 sit at fp32 summation-order level;
    the north star asks for 1e-3)."""
    rng = np.random.RandomState(V + d)
    if kind == "bert":
        model = rbm_b200.model_factory(bert_args(V, Ln, d, nb, h, seed=2))
        model.load_state_dict(ob.random_state_dict(V, Ln, d, nb, seed=5))
        model.to(DEV).train()
        tok = rng.randint(1, V + 1, size=(B, Ln)).astype(np.int64)
        tok[0, : Ln // 3] = 0
        lab = np.where((rng.rand(B, Ln) < 0.2) & (tok != 0), tok, 0)
        tok = np.where(lab != 0, V + 1, tok)
        batch = (torch.from_numpy(tok), torch.from_numpy(lab))
        loss = model.loss(*batch)
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
        ref = ob.loss(sd, batch[0], batch[1], nb, h)
    else:
        model = rbm_b200.model_factory(sas_args(V, Ln, d, nb, h))
        model.load_state_dict(osr.random_state_dict(V, Ln, d, nb, seed=5))
        model.to(DEV).train()
        seq = rng.randint(1, V + 1, size=(B, Ln)).astype(np.int64)
        for b in range(B):
            seq[b, : rng.randint(0, Ln - 2)] = 0  # left padding of varying length (Amazon-Beauty-like short histories)
        pos = np.where(seq != 0, rng.randint(1, V + 1, size=(B, Ln)), 0)
        neg = np.where(seq != 0, rng.randint(1, V + 1, size=(B, Ln)), 0)
        loss = model.loss(seq, pos, neg)
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
        ref = osr.loss(sd, torch.from_numpy(seq), torch.from_numpy(pos), torch.from_numpy(neg), nb, h)
    loss.backward()
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    for k, prm in model.named_parameters():
        gr = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
        floor = 1e-4
        if k.endswith("attention.linear_layers.1.bias"):
            # the key-projection bias gradient is exactly zero in exact arithmetic (softmax is shift-invariant): both sides
            # hold only rounding noise there, which is judged against (1 % of) the scale of the block's query-bias gradient
            floor = 10.0 * float(sd[k.replace("linear_layers.1.bias", "linear_layers.0.bias")].grad.abs().max())
        relclose(prm.grad.cpu().numpy(), gr.numpy(), 1e-3, floor=floor, msg=k)


@pytest.mark.parametrize("kind", ["bert", "sas"])
def test_dataloader_factory_end_to_end(kind, tmp_path):
    """dataloader_factory -> (train, val, test) device loaders -> trainer loops' hooks: text file in, the reference's batch wire
    formats out (checked against the oracle's eval batches / negatives), one training step and one validation batch run."""
    from oracle import batches as obt
    from rbm_b200.dataloaders import data_partition
    rs = np.random.RandomState(8)
    V, Ln = 90, 12
    lines = ["%d %d" % (u, i) for u in range(1, 41) for i in rs.randint(1, V + 1, size=rs.randint(3, 30))]
    f = tmp_path / "toy.txt"
    f.write_text("\n".join(lines) + "\n")
    common = dict(model_code=kind, data_path=str(f), data_name="toy.txt", max_len=Ln, prop_sliding_window=0.3, device=DEV,
                  load_processed_dataset=False, dataloader_random_seed=3, worker_number=0, train_batch_size=16, val_batch_size=8,
                  test_batch_size=8, test_negative_sampler_code="random", test_negative_sample_size=20,
                  test_negative_sampling_seed=98765, bert_mask_prob=0.3, optimizer="Adam", lr=1e-3, weight_decay=0, momentum=None,
                  decay_step=10, gamma=1.0, num_epochs=1, metric_ks=[1, 5, 10], best_metric="NDCG@10", resume_path=None, l2_emb=0.0)
    extra = vars(bert_args(0, Ln, 16, 1, 2, p=0.1, seed=1)) if kind == "bert" else vars(sas_args(0, Ln, 16, 1, 1, p=0.1))
    args = SimpleNamespace(**{**extra, **common})
    train, val, test = rbm_b200.dataloader_factory(args)
    tr, va, te, n, itemnum = data_partition(str(f), Ln, 0.3)
    assert args.num_items == itemnum and len(val) == (n + 7) // 8
    # evaluation batches == the oracle's, with the oracle's negatives (bit-exact contract)
    seen = [sorted(set(tr[u]) | set(va[u]) | set(te[u])) for u in range(n)]
    negs = obt.negative_samples(seen, itemnum, 20, 98765, ops.NEG_SITE)
    mask = itemnum + 1 if kind == "bert" else -1
    for loader, hist, ans in ((val, tr, [v[0] for v in va]), (test, [tr[u] + va[u] for u in range(n)], [t[0] for t in te])):
        got = [torch.cat([b[j] for b in loader]).cpu().numpy() for j in range(3)]
        ref = obt.eval_batch(hist, ans, negs, list(range(n)), Ln, mask)
        for g_, r_ in zip(got, ref):
            np.testing.assert_array_equal(g_, r_)
    model = rbm_b200.model_factory(args)
    trainer = rbm_b200.trainer_factory(args, model, train, val, test, None)
    model.train()
    batch = next(iter(train))
    assert all(t.shape == (16, Ln) and t.dtype == torch.int64 for t in batch)
    l0 = trainer.train_step(batch).item()
    assert np.isfinite(l0)
    model.eval()
    with torch.no_grad():
        m = trainer.calculate_metrics(next(iter(val)))
    assert set(m) == {"%s@%d" % (nm, k) for nm in ("Recall", "NDCG", "MRR") for k in (1, 5, 10)}
    assert all(0.0 <= v <= 1.0 for v in m.values())


@pytest.mark.parametrize("name", ["bert_cfg4s", "sas_cfg3s"])
def test_wide_models_vs_reference_golden(name):
    """The d = 256 / 128 model shapes of BASELINE configs[3] / configs[2] against numbers produced by the UNMODIFIED reference
    (tests/golden/{bert_cfg4s,sas_cfg3s}.npz): loss within 3e-4 relative (north star: 1e-3), per-parameter gradient norms within
    3e-3 (absolute floor 1e-5 of the largest norm: the key-projection bias gradient is pure rounding noise on both sides)."""
    z = load(name)
    V, Ln, d, nb, h, B, seed = z["cfg"].tolist()
    if name.startswith("bert"):
        model = rbm_b200.model_factory(bert_args(V, Ln, d, nb, h, seed=seed))
        model.load_state_dict(ob.random_state_dict(V, Ln, d, nb, seed=seed))
        model.to(DEV).train()
        loss = model.loss(torch.from_numpy(z["tokens"]), torch.from_numpy(z["labels"]))
    else:
        model = rbm_b200.model_factory(sas_args(V, Ln, d, nb, h))
        model.load_state_dict(osr.random_state_dict(V, Ln, d, nb, seed=seed))
        model.to(DEV).train()
        loss = model.loss(z["seq"], z["pos"], z["neg"])
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 3e-4 * abs(float(z["loss"])), (loss.item(), float(z["loss"]))
    gn = np.array([p.grad.norm().item() for _, p in model.named_parameters()])
    np.testing.assert_allclose(gn, z["grad_norms"], rtol=3e-3, atol=1e-5 * float(z["grad_norms"].max()))


def test_bert_cfg4_million_items_vs_oracle():
    """BASELINE configs[3] at its real catalogue size: BERT4Rec d = 256, h = 4, 10^6 items (2 blocks, B = 4 keeps the CPU oracle
    to seconds) -- loss and the gradients of the output layer, the token table and a dense weight against the chunked CPU
    restatement `oracle.bert4rec.loss_masked_only` (the reference itself cannot allocate 800 MB of logits per
    sequence).  Tolerances: loss 1e-4 relative, gradients 1e-3 of their scale."""
    V, Ln, d, nb, h, B = 1_000_000, 200, 256, 2, 4, 4
    sd = ob.random_state_dict(V, Ln, d, nb, seed=11)
    a = bert_args(V, Ln, d, nb, h, p=0.0, seed=11)
    with torch.device(DEV):
        model = rbm_b200.model_factory(a)
    model = model.to(DEV)
    model.load_state_dict(sd)
    model.train()
    rs = np.random.RandomState(2)
    tok = rs.randint(1, V + 1, size=(B, Ln)).astype(np.int64)
    tok[1, :40] = 0
    lab = np.where((rs.rand(B, Ln) < 0.15) & (tok != 0), tok, 0)
    tokm = np.where(lab != 0, V + 1, tok)
    loss = model.loss(torch.from_numpy(tokm), torch.from_numpy(lab))
    loss.backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = ob.loss_masked_only(leaves, torch.from_numpy(tokm), torch.from_numpy(lab), nb, h)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    got = dict(model.named_parameters())
    for k in ("out.weight", "out.bias", "bert.embedding.token.weight", "bert.transformer_blocks.1.feed_forward.w_2.weight",
              "bert.transformer_blocks.0.attention.linear_layers.0.weight"):
        g_ref = leaves[k].grad
        err = float((got[k].grad.cpu() - g_ref).abs().max())
        assert err <= 1e-3 * float(g_ref.abs().max()), (k, err, float(g_ref.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,V,Ln,d,nb,h,B", [("sas", 500, 50, 64, 2, 1, 37), ("sas", 700, 50, 128, 2, 2, 9), ("bert", 300, 40, 64, 2, 2, 21),
                                                 ("bert", 200, 200, 256, 1, 4, 3), ("sas", 90, 7, 16, 1, 2, 4)])
def test_eval_last_position_only(kind, V, Ln, d, nb, h, B):
    """K20: in evaluation the final block runs for the last position alone (rbm_attn_last_query + [B, d] projections).  Its hidden
    state, the candidate scores and the full-catalogue top-10 must match the all-positions computation the reference performs
    (NN/trainers/bert.py:47, NN/models/sas_model/sas.py:107-118): hidden 1e-5 of scale, ids identical on clear-gap rows."""
    rng = np.random.RandomState(V + Ln)
    if kind == "bert":
        model = rbm_b200.model_factory(bert_args(V, Ln, d, nb, h, seed=3))
        model.load_state_dict(ob.random_state_dict(V, Ln, d, nb, seed=8))
    else:
        model = rbm_b200.model_factory(sas_args(V, Ln, d, nb, h))
        model.load_state_dict(osr.random_state_dict(V, Ln, d, nb, seed=8))
    model.to(DEV).eval()
    seq = rng.randint(1, V + 1, size=(B, Ln)).astype(np.int64)
    for b in range(B):
        seq[b, : rng.randint(0, Ln - 1)] = 0  # left padding
    seq[0, :] = 0  # a sequence of padding only (empty history)
    if kind == "bert":
        seq[:, -1] = V + 1  # the mask token the evaluation batches end with
    x = torch.from_numpy(seq).to(DEV)
    with torch.no_grad():
        full = (model.hidden_states(x) if kind == "bert" else model.log2feats(x))[:, -1, :]
        last = model.last_hidden(x)
        assert last.shape == full.shape
        relclose(last.cpu().numpy(), full.cpu().numpy(), 1e-5, msg="last hidden")
        if kind == "sas":  # SASRec: the blocks before the last one on the non-padding rows only (forced) / on every row (forced)
            for frac in (1.0, -1.0):
                model.LIVE_ROWS_MAX_FRACTION = frac
                relclose(model.last_hidden(x).cpu().numpy(), full.cpu().numpy(), 1e-5, msg="last hidden, live-row share limit %g" % frac)
            model.LIVE_ROWS_MAX_FRACTION = 1.0
        cand = torch.from_numpy(rng.randint(1, V + 1, size=(B, 11)).astype(np.int64)).to(DEV)
        sc = model.candidate_scores(x, cand) if kind == "bert" else model.predict(x, cand)
        w = model.out.weight if kind == "bert" else model.sas.item_emb.weight
        ref = (full.double().unsqueeze(1) * w[cand].double()).sum(-1)
        if kind == "bert":
            ref = ref + model.out.bias[cand].double()
        relclose(sc.cpu().numpy(), ref.cpu().numpy(), 1e-5, msg="candidate scores")
    # the training-mode / grad-enabled path still computes every position
    assert model.last_hidden(x).shape == full.shape


@pytest.mark.gpu
def test_rows_gather_scatter():
    """csrc/rows.cu against torch indexing: gather (zero rows past the count), scatter with / without a fill vector, and the
    gradients (scatter backward = gather + the fixed-order column sum over the padding rows)."""
    g = torch.Generator(device=DEV).manual_seed(3)
    for n, d, frac in ((1000, 64, 0.2), (4097, 128, 0.9), (300, 16, 0.0), (513, 256, 1.0), (700, 100, 0.5), (260, 1024, 0.3)):
        tok = (torch.rand(n, device=DEV, generator=g) < frac).long() * torch.randint(1, 50, (n,), device=DEV, generator=g)
        cnt = int((tok != 0).sum())
        cap = max(128, -(-cnt // 128) * 128) + 128
        live = ops.LiveRows(tok, cap)
        assert int(live.count.item()) == cnt
        idx = torch.nonzero(tok).flatten()
        np.testing.assert_array_equal(live.rows[:cnt].cpu().numpy(), idx.cpu().numpy())
        x = torch.randn(n, d, device=DEV, generator=g, requires_grad=True)
        fill = torch.randn(d, device=DEV, generator=g, requires_grad=True)
        xc = ops.rows_gather(x, live)
        assert xc.shape == (cap, d) and torch.equal(xc[:cnt], x[idx]) and not xc[cnt:].any()
        y = ops.rows_scatter(xc * 2.0, fill, live)
        ref = torch.where((tok != 0)[:, None], x * 2.0, fill[None, :].expand(n, d))
        assert torch.equal(y, ref)
        dy = torch.randn(n, d, device=DEV, generator=g)
        y.backward(dy)
        gx = torch.where((tok != 0)[:, None], dy * 2.0, torch.zeros_like(dy))
        assert torch.equal(x.grad, gx)
        gf = dy[tok == 0].double().sum(0)
        relclose(fill.grad.cpu().numpy(), gf.cpu().numpy(), 1e-5, msg="fill grad")
        y0 = ops.rows_scatter(xc.detach(), None, live)
        assert torch.equal(y0, torch.where((tok != 0)[:, None], x.detach(), torch.zeros_like(x)))


@pytest.mark.gpu
@pytest.mark.parametrize("V,Ln,d,nb,h,B,p", [(300, 50, 128, 2, 2, 64, 0.0), (90, 12, 32, 2, 1, 9, 0.25), (500, 50, 64, 1, 1, 300, 0.2),
                                             (120, 80, 32, 1, 2, 10, 0.2), (200, 64, 256, 1, 2, 7, 0.1), (60, 33, 48, 2, 2, 5, 0.3)])
def test_sas_live_rows_training(V, Ln, d, nb, h, B, p):
    """SASRec with the token-wise layers on the non-padding rows only (Amazon-Beauty-like left padding: 80-90 % of the positions):
    loss and EVERY gradient against the oracle -- dropout on: the oracle consumes the Philox masks the kernels used, the
    element-wise sites of this path being indexed by (live-row ordinal, column) -- and against the dense path at p = 0."""
    from oracle.common import DropoutPlan
    if os.environ.get("RBM_SAS_LIVE_ROWS", "1") == "0":
        pytest.skip("the live-row path is switched off by the environment")
    rng = np.random.RandomState(V + B)
    seq = rng.randint(1, V + 1, size=(B, Ln)).astype(np.int64)
    for b in range(B):
        seq[b, : Ln - rng.randint(0, max(2, Ln // 5))] = 0  # 0 .. L/5 - 1 items, right-aligned; some rows are empty
    seq[B // 2] = rng.randint(1, V + 1, size=Ln)  # one full-length history
    pos = np.where(seq != 0, rng.randint(1, V + 1, size=(B, Ln)), 0)
    neg = np.where(seq != 0, rng.randint(1, V + 1, size=(B, Ln)), 0)
    pos[0, 0] = 3  # a labelled position on a padding row: its features are the last LayerNorm's beta
    neg[0, 0] = 5
    model = rbm_b200.model_factory(sas_args(V, Ln, d, nb, h, p=p))
    model.load_state_dict(osr.random_state_dict(V, Ln, d, nb, seed=11))
    model.to(DEV).train()
    model.dropout_seed = 777
    assert model._live_rows(torch.from_numpy(seq).to(DEV)) is not None
    step0 = model._step
    launches0 = rbm_b200.lib.launch_count
    loss = model.loss(seq, pos, neg)
    loss.backward()
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    tseq, tpos, tneg = (torch.from_numpy(x) for x in (seq, pos, neg))
    if p == 0.0:
        ref = osr.loss(sd, tseq, tpos, tneg, nb, h)
    else:
        live = np.flatnonzero(seq.reshape(-1))
        cap = model._capacity(max(len(live), int((pos != 0).sum()), int((neg != 0).sum())))
        base = step0 * 64
        masks = {0: ops.dropout_mask(B * Ln * d, p, 777, base, DEV).cpu()}
        for b in range(nb):
            s = osr.block_sites(b)
            masks[s["attn"]] = ops.dropout_mask_attn(B * h * Ln, Ln, p, 777, base + s["attn"], DEV).cpu()
            for key in ("ffn1", "ffn2"):
                mc = ops.dropout_mask(cap * d, p, 777, base + s[key], DEV).cpu().reshape(cap, d)
                full = torch.ones(B * Ln, d, dtype=mc.dtype)  # padding rows: whatever, they are multiplied by zero
                full[torch.from_numpy(live)] = mc[: len(live)]
                masks[s[key]] = full
        ref = osr.loss(sd, tseq, tpos, tneg, nb, h, p=p, drop=DropoutPlan(True, masks))
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()), (loss.item(), ref.item())
    for k, prm in model.named_parameters():
        gr = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
        relclose(prm.grad.cpu().numpy(), gr.numpy(), 1e-3, msg=k)
    if p == 0.0:  # the dense path on the same batch: same loss to rounding
        model.zero_grad()
        model.LIVE_ROWS_MAX_FRACTION = -1.0
        dense = model.loss(seq, pos, neg)
        assert abs(dense.item() - loss.item()) < 1e-6 * abs(loss.item())
    del launches0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sas", "bert"])
def test_live_rows_cuda_graph(kind):
    """The captured step with a fixed row capacity (SASRec: non-padding rows; BERT4Rec: labelled rows of the final block): replays
    are bit-identical to eager steps run with the same capacity; a batch with more such rows than the capacity takes an eager
    step and training continues."""
    if os.environ.get("RBM_SAS_LIVE_ROWS" if kind == "sas" else "RBM_BERT_LABEL_ROWS", "1") == "0":
        pytest.skip("the row-compacting path is switched off by the environment")
    V, Ln, d, B = 80, 20, 32, 48
    rs = np.random.RandomState(5)

    def batch(i, keep=4, p_label=0.15):
        if kind == "sas":
            s = rs.randint(1, V + 1, size=(B, Ln))
            for b in range(B):
                s[b, : Ln - rs.randint(1, keep + 1)] = 0
            p_ = np.where(s != 0, rs.randint(1, V + 1, size=(B, Ln)), 0)
            n_ = np.where(s != 0, rs.randint(1, V + 1, size=(B, Ln)), 0)
            return tuple(torch.from_numpy(x).to(DEV) for x in (s, p_, n_))
        t = rs.randint(1, V + 1, size=(B, Ln)); t[:, : i % 3] = 0
        l = np.where((rs.rand(B, Ln) < p_label) & (t != 0), t, 0)
        return torch.from_numpy(np.where(l != 0, V + 1, t)).to(DEV), torch.from_numpy(l).to(DEV)

    a = sas_args(V, Ln, d, 2, 1, p=0.2) if kind == "sas" else bert_args(V, Ln, d, 2, 2, p=0.2, seed=3)
    common = dict(optimizer="Adam", lr=2e-3, weight_decay=0, momentum=None, decay_step=50, gamma=1.0, num_epochs=1, metric_ks=[10],
                  best_metric="NDCG@10", train_batch_size=B, resume_path=None, **({"l2_emb": 0.0} if kind == "sas" else {}))
    batches = [batch(i) for i in range(5)]
    torch.manual_seed(0)
    m1, m2 = rbm_b200.model_factory(a), rbm_b200.model_factory(a)
    m2.load_state_dict(m1.state_dict())
    m2.dropout_seed = m1.dropout_seed
    t1 = rbm_b200.trainer_factory(SimpleNamespace(**vars(a), **common), m1, None, None, None, None)
    t2 = rbm_b200.trainer_factory(SimpleNamespace(**vars(a), **common), m2, None, None, None, None)
    m1.train(); m2.train()
    t2.capture_train_step(batches[0])
    cap = t2._graph_row_cap
    assert 0 < cap < B * Ln
    # a batch that does not fit the capacity: eager step inside the graphed trainer, dense / self-sized in the twin
    if kind == "sas":
        sf = rs.randint(1, V + 1, size=(B, Ln))
        big = tuple(torch.from_numpy(x).to(DEV) for x in (sf, rs.randint(1, V + 1, size=(B, Ln)), rs.randint(1, V + 1, size=(B, Ln))))
    else:
        big = batch(1, p_label=0.45)
    assert m2.live_row_count(*big) > cap
    # (the twin takes all its steps first: the device-side step counter the captured trainer installs is process-wide)
    m1._row_cap = cap  # the eager twin runs with the captured capacity ...
    eager = [t1.train_step(b).item() for b in batches]
    m1._row_cap = None  # ... and sizes the oversized batch by itself, like the eager step inside the graphed trainer
    eager.append(t1.train_step(big).item())
    m1._row_cap = cap
    eager.append(t1.train_step(batches[1]).item())
    graphed = [t2.train_step(b).item() for b in batches + [big, batches[1]]]
    assert eager == graphed, (eager, graphed)
    t2.release_train_graph()
    for (k, v1), (_, v2) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(v1, v2), k
