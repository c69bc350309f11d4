"""CPU-only checks: the C-ABI library builds/loads and exports every symbol include/rbm.h declares, the host side
mirrors the reference's API (registries, state_dict names/shapes, bit-identical initialisation), inputs producers
emit the reference's wire formats, and nothing falls back to the CPU."""
import ctypes
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import __graft_entry__ as entry

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="session", autouse=True)
def built():
    entry.build()  # nvcc cross-compiles for sm_100a without a GPU; no-op when up to date


def load(name):
    z = np.load(os.path.join(G, name + ".npz"))
    return {k: z[k] for k in z.files}


def test_library_exports_every_declared_symbol():
    import rbm_b200
    from rbm_b200 import lib as L
    declared = L.header_functions()
    assert len(declared) >= 38
    assert sorted(L.SIGNATURES) == declared  # the binding table covers the header exactly
    cdll = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(cdll, name), name
    lib = L.load()
    assert lib.rbm_abi_version() == 1
    assert isinstance(lib.rbm_last_error(), bytes)
    # no torch / python symbols leak through the boundary: plain C ABI
    out = os.popen("nm -D --defined-only '%s'" % L.LIB_PATH).read()
    exported = set(re.findall(r" T (\w+)", out))
    assert set(declared) <= exported
    assert not [s for s in exported if "torch" in s.lower() or "at::" in s]


def test_size_queries_without_gpu():
    from rbm_b200 import lib as L
    lib = L.load()
    assert lib.rbm_scatter_ws_bytes(1000, 50) > 4 * 1000 * 4
    assert lib.rbm_linear_bwd_weight_ws_bytes(4096, 64, 64) >= 64 * 64 * 4
    assert lib.rbm_ce_ws_bytes(1000, 3417, 64) > 3417 * 64 * 4
    assert lib.rbm_attn_bwd_ws_bytes(2, 50, 2) == 2 * 50 * 2 * 4
    assert lib.rbm_score_topk_ws_bytes(100, 5000, 10) > 0


def test_argument_errors_are_reported_not_hidden():
    from rbm_b200 import lib as L
    lib = L.load()
    # invalid arguments are rejected before any launch: works without a GPU
    rc = lib.rbm_embed_fwd(None, None, None, None, 8, 4, 6, 10, 1.0, 0, 0.0, 0, 0, None)
    assert rc < 0 and b"null pointer" in lib.rbm_last_error()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.rbm_layernorm_fwd(p, p, p, p, p, 4, 6, 1e-6, 0, None)
    assert rc < 0 and b"unsupported d=6" in lib.rbm_last_error()
    rc = lib.rbm_attn_fwd(p, 4, p, 4, p, 4, None, p, 4, None, 1, 300, 1, 8, 1, 1.0, 0.0, 0, 0, None)
    assert rc < 0 and b"L<=256" in lib.rbm_last_error()
    rc = lib.rbm_topk_rows(p, 8, p, p, 2, 8, 40, 0, None)
    assert rc < 0
    with pytest.raises(RuntimeError):
        L.check(rc, "topk_rows")


def test_no_cpu_fallback():
    import rbm_b200
    from rbm_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.layernorm(torch.randn(4, 8), torch.ones(8), torch.zeros(8), 1e-6, 1)
    with pytest.raises(RuntimeError, match="GPU"):
        rbm_b200.recalls_ndcgs_and_mrr_for_ks(torch.randn(4, 8), torch.zeros(4, 8, dtype=torch.long), [1, 5])
    src = "".join(open(os.path.join(os.path.dirname(rbm_b200.lib.__file__), f)).read()
                  for f in ("ops.py", "lib.py", "optim.py", "dist.py", "models/bert.py", "models/sas.py", "trainers/utils.py"))
    assert "oracle" not in src  # the product never imports the checker


def bert_args(V, Ln, d, nb, h, seed):
    return SimpleNamespace(model_code="bert", num_items=V, max_len=Ln, device="cpu", model_init_seed=seed, bert_num_blocks=nb,
                           bert_num_heads=h, bert_hidden_units=d, bert_dropout=0.1, bert_hidden_dropout=0.1)


def sas_args(V, Ln, d, nb, h):
    return SimpleNamespace(model_code="sas", num_items=V, max_len=Ln, device="cpu", sas_hidden_units=d, sas_num_blocks=nb,
                           sas_heads=h, sas_dropout=0.2)


def test_registries_and_state_dict_match_the_reference():
    import rbm_b200
    assert set(rbm_b200.MODELS) == {"bert", "sas"} and set(rbm_b200.TRAINERS) == {"bert", "sas"}
    for name in ("bert_tiny", "bert_odd"):
        z = load(name)
        V, Ln, d, nb, h, B, seed = z["cfg"].tolist()
        m = rbm_b200.model_factory(bert_args(V, Ln, d, nb, h, seed))
        sd = m.state_dict()
        ref = {k[3:]: v for k, v in z.items() if k.startswith("sd.")}
        assert list(sd.keys()) == list(ref.keys())  # same names, same order
        for k in ref:
            assert tuple(sd[k].shape) == ref[k].shape, k
            # same torch layers created in the same order after fix_random_seed_as(model_init_seed): same bits
            np.testing.assert_array_equal(sd[k].numpy(), ref[k], err_msg=k)
    for name in ("sas_tiny", "sas_odd"):
        z = load(name)
        V, Ln, d, nb, h, B, seed = z["cfg"].tolist()
        torch.manual_seed(seed)
        m = rbm_b200.model_factory(sas_args(V, Ln, d, nb, h))
        sd = m.state_dict()
        ref = {k[3:]: v for k, v in z.items() if k.startswith("sd.")}
        assert list(sd.keys()) == list(ref.keys())
        for k in ref:
            np.testing.assert_array_equal(sd[k].numpy(), ref[k], err_msg=k)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in ref.items()})  # reference checkpoints load


def test_fused_adam_state_dict_layout():
    import rbm_b200
    p = torch.nn.Parameter(torch.zeros(3))
    opt = rbm_b200.FusedAdam([p], lr=1e-3)
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(3))], lr=1e-3)
    assert opt.state_dict()["param_groups"][0].keys() >= {"lr", "betas", "eps", "weight_decay", "params"}
    sch = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)  # NN/trainers/base.py:40
    opt.step()  # no grads: nothing to do, must not touch the GPU
    sch.step()
    assert abs(opt.param_groups[0]["lr"] - 5e-4) < 1e-12
    assert ref.defaults["betas"] == opt.defaults["betas"] and ref.defaults["eps"] == opt.defaults["eps"]


def test_wire_formats_of_the_batch_builders():
    from rbm_b200.dataloaders import (synthetic_interactions, sliding_window_partition, BertBatcher, SasBatcher, eval_sequences,
                                      uniform_negative_candidates)
    hist = synthetic_interactions(300, 500, 60, 5, seed=1)
    assert all(h.min() >= 1 and h.max() <= 500 for h in hist)
    assert all((h[1:] != h[:-1]).all() for h in hist)
    train, valid, test, n, V = sliding_window_partition(hist, 50, 0.3)
    assert n == len(train) == len(valid) == len(test) and V <= 500
    assert all(len(t) <= 48 and len(v) == 1 and len(s) == 1 for t, v, s in zip(train, valid, test))
    # data_partition's window rule (NN/dataloaders/__init__.py:41-53) on one long history
    long = [np.arange(1, 131)]
    tr, va, te, n2, _ = sliding_window_partition(long, 50, 0.3)
    starts = list(range(130 - 50, 0, -15))[::-1]
    assert n2 == len(starts) and [t[0] for t in tr] == [s + 1 for s in starts]
    assert all(t == list(range(s + 1, s + 49)) and v == [s + 49] and e == [s + 50] for t, v, e, s in zip(tr, va, te, starts))

    bb = BertBatcher(train, V, 50, 0.3, seed=0)
    tok, lab = bb.batch(256)
    assert tok.dtype == np.int64 and tok.shape == lab.shape == (256, 50)
    pad = tok == 0
    assert (pad[:, :-1] >= pad[:, 1:]).all()  # left padding only
    assert ((lab != 0) <= ~pad).all() and tok.max() <= V + 1
    masked = lab != 0
    assert 0.25 < masked.sum() / (~pad).sum() < 0.35
    frac_mask = (tok[masked] == V + 1).mean()
    assert 0.74 < frac_mask < 0.86  # 80 % [MASK], 10 % random, 10 % kept
    assert (tok[~masked & ~pad] != V + 1).all()

    sb = SasBatcher(train, V, 50, seed=0)
    seq, pos, neg = sb.batch(64)
    assert seq.shape == pos.shape == neg.shape == (64, 50) and seq.dtype == np.int64
    assert ((seq == 0) == (pos == 0)).all()
    inner = (seq[:, 1:] != 0) & (seq[:, :-1] != 0)
    assert (pos[:, :-1][inner] == seq[:, 1:][inner]).all()  # pos is seq shifted by one
    for b in range(64):
        real = pos[b] != 0
        assert not np.isin(neg[b][real], np.concatenate([seq[b][seq[b] != 0], pos[b][real]])).any()
        assert (neg[b][~real] == 0).all()

    ev = eval_sequences(train, valid, 50, mask_token=V + 1)
    assert ev.shape == (n, 50) and (ev[:, -1] == V + 1).all() and (ev[:, -2] == [v[0] for v in valid]).all()
    cands, labels = uniform_negative_candidates(test, V, 100)
    assert cands.shape == (n, 101) and (labels[:, 0] == 1).all() and labels[:, 1:].sum() == 0
    assert (cands[:, 1:] != cands[:, :1]).all() and cands.min() >= 1 and cands.max() <= V


def test_registries_mirror_the_reference():
    """MODELS / TRAINERS / DATALOADERS keys and code() classmethods (NN/models/__init__.py:4-12, NN/trainers/__init__.py:4-12,
    NN/dataloaders/__init__.py:11-14); an unknown model code is an error, as in the reference's dict lookup."""
    import rbm_b200
    assert set(rbm_b200.MODELS) == set(rbm_b200.TRAINERS) == set(rbm_b200.DATALOADERS) == {"bert", "sas"}
    for reg in (rbm_b200.MODELS, rbm_b200.TRAINERS, rbm_b200.DATALOADERS):
        for code, cls in reg.items():
            assert cls.code() == code
    from types import SimpleNamespace
    with pytest.raises(KeyError):
        rbm_b200.dataloader_factory(SimpleNamespace(model_code="gru", max_len=5), dataset=[[], [], [], 0, 0])


def test_row_capacity_rules_host_side():
    """Host logic of the row-compacting training paths (models/sas.py live rows, models/bert.py labelled rows): one capacity rule
    for eager and captured steps, dense path above the live / labelled share, and the per-batch fit check a captured step relies
    on -- incl. the per-sequence query capacity of BERT4Rec's final block.  CPU tensors only."""
    from types import SimpleNamespace
    import rbm_b200
    from rbm_b200.models.bert import BERTModel
    from rbm_b200.models.sas import SASModel
    # the rules: headroom, whole tiles, monotone
    assert SASModel._capacity(0) == 128 and SASModel._capacity(1000) == 1536 and SASModel._capacity(28672) % 128 == 0
    caps = [SASModel._capacity(c) for c in range(0, 5000, 37)]
    assert all(a <= b for a, b in zip(caps, caps[1:])) and all(c >= n + n // 8 for c, n in zip(caps, range(0, 5000, 37)))
    cap, lq = BERTModel._capacities(30000, 47)
    assert cap % 128 == 0 and cap >= 33750 + 8 * 173 and lq % 16 == 0 and lq >= 58
    # SASRec: capacity from the largest non-zero count of (seq, pos, neg); dense path above 60 % live rows
    sas = rbm_b200.model_factory(SimpleNamespace(model_code="sas", num_items=50, max_len=20, device="cpu", sas_hidden_units=16, sas_num_blocks=1,
                                                 sas_heads=1, sas_dropout=0.1))
    seq = torch.zeros(64, 20, dtype=torch.int64)
    seq[:, -3:] = 7
    pos = seq.clone()
    pos[0, 0] = 9  # one label on a padding row
    assert sas.live_row_count(seq, pos, seq) == 64 * 3 + 1
    assert sas.row_capacity_for(seq, pos, seq) == SASModel._capacity(64 * 3 + 1)
    assert sas.row_capacity_for(torch.ones(64, 20, dtype=torch.int64)) == 0
    # BERT4Rec: rows from the labels, per-sequence query capacity remembered for the fit check
    bert = rbm_b200.model_factory(SimpleNamespace(model_code="bert", num_items=50, max_len=40, device="cpu", model_init_seed=0, bert_num_blocks=1,
                                                  bert_num_heads=2, bert_hidden_units=64, bert_dropout=0.1, bert_hidden_dropout=0.1))
    tok = torch.ones(32, 40, dtype=torch.int64)
    lab = torch.zeros(32, 40, dtype=torch.int64)
    lab[:, :5] = 3
    assert bert._label_counts(lab) == (160, 5)
    cap = bert.row_capacity_for(tok, lab)
    assert cap == min(BERTModel._capacities(160, 5)[0], 1280) and bert._graph_lq == 16
    assert bert.live_row_count(tok, lab) == 160
    lab2 = lab.clone()
    lab2[3, :20] = 3  # one sequence with more labels than the captured per-sequence query capacity
    assert bert.live_row_count(tok, lab2) > (1 << 40)
    assert bert.row_capacity_for(tok, torch.ones(32, 40, dtype=torch.int64)) == 0 and bert._graph_lq is None
