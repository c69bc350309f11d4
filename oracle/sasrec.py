"""SASRec CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Functional fp32 restatement over a reference-shaped ``state_dict`` (keys ``sas.item_emb.weight``
... ``sas.last_layernorm.bias``).  ``NN/`` =
``/root/reference/NerualNetwork/bert4rec&sas4rec/``.  ``nn.MultiheadAttention``'s arithmetic
lives in torch (``F.multi_head_attention_forward``, ``need_weights=True`` branch); it is restated
explicitly here and validated against the installed torch by tests/golden/make_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .common import DropoutPlan, NO_DROPOUT, torch_layernorm

SITE_EMB = 0
LN_EPS = 1e-8  # NN/models/sas_model/sas.py:39,42,50


def block_sites(b: int):
    base = 1 + 3 * b
    return dict(attn=base, ffn1=base + 1, ffn2=base + 2)


def num_sites(n_blocks: int) -> int:
    return 1 + 3 * n_blocks


def causal_mha(sd, pfx: str, query: torch.Tensor, keyval: torch.Tensor, heads: int, p: float,
               drop: DropoutPlan, site: int) -> torch.Tensor:
    """``nn.MultiheadAttention(d, h, p)(Q, seqs, seqs, attn_mask=~tril)`` as called at
    NN/models/sas_model/sas.py:75-76, batch-first.  Packed in-proj: q from ``query`` (the
    LayerNorm-ed stream), k/v from ``keyval`` (the UN-normalised stream); q is scaled by
    1/sqrt(d_h) before q.k^T; masked (future) scores are -inf; no key-padding mask (commented
    out at sas.py:77)."""
    B, L, d = query.shape
    dh = d // heads
    w, bias = sd[f"{pfx}.in_proj_weight"], sd[f"{pfx}.in_proj_bias"]
    q = F.linear(query, w[:d], bias[:d])
    k = F.linear(keyval, w[d:2 * d], bias[d:2 * d])
    v = F.linear(keyval, w[2 * d:], bias[2 * d:])
    q = q.view(B, L, heads, dh).transpose(1, 2) * math.sqrt(1.0 / dh)
    k = k.view(B, L, heads, dh).transpose(1, 2)
    v = v.view(B, L, heads, dh).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-2, -1))
    future = ~torch.tril(torch.ones(L, L, dtype=torch.bool))
    scores = scores.masked_fill(future, float("-inf"))
    pr = drop(F.softmax(scores, dim=-1), p, site)
    ctx = torch.matmul(pr, v).transpose(1, 2).contiguous().view(B, L, d)
    return F.linear(ctx, sd[f"{pfx}.out_proj.weight"], sd[f"{pfx}.out_proj.bias"])


def point_wise_ffn(sd, pfx: str, x: torch.Tensor, p: float, drop: DropoutPlan, s1: int, s2: int) -> torch.Tensor:
    """``PointWiseFeedForward.forward`` NN/models/sas_model/sas.py:16-20: Conv1d(d,d,1) == per-token
    Linear; conv1 -> dropout1 -> relu -> conv2 -> dropout2, ``+= inputs``."""
    w1, b1 = sd[f"{pfx}.conv1.weight"].squeeze(-1), sd[f"{pfx}.conv1.bias"]
    w2, b2 = sd[f"{pfx}.conv2.weight"].squeeze(-1), sd[f"{pfx}.conv2.bias"]
    u = torch.relu(drop(F.linear(x, w1, b1), p, s1))
    return x + drop(F.linear(u, w2, b2), p, s2)


def log2feats(sd: Dict[str, torch.Tensor], seq: torch.Tensor, n_blocks: int, heads: int, p: float = 0.0,
              drop: DropoutPlan = NO_DROPOUT) -> torch.Tensor:
    """``SAS.log2feats`` NN/models/sas_model/sas.py:59-88 -> [B, L, d]."""
    d = sd["sas.item_emb.weight"].size(1)
    L = seq.size(1)
    x = F.embedding(seq, sd["sas.item_emb.weight"], padding_idx=0) * (d ** 0.5)
    x = x + sd["sas.pos_emb.weight"][:L].unsqueeze(0)
    x = drop(x, p, SITE_EMB)
    keep = (seq != 0).unsqueeze(-1)
    x = x * keep
    for b in range(n_blocks):
        s = block_sites(b)
        Q = torch_layernorm(x, sd[f"sas.attention_layernorms.{b}.weight"], sd[f"sas.attention_layernorms.{b}.bias"], LN_EPS)
        a = causal_mha(sd, f"sas.attention_layers.{b}", Q, x, heads, p, drop, s["attn"])
        x = Q + a  # residual on the NORMALISED query, sas.py:79
        x = torch_layernorm(x, sd[f"sas.forward_layernorms.{b}.weight"], sd[f"sas.forward_layernorms.{b}.bias"], LN_EPS)
        x = point_wise_ffn(sd, f"sas.forward_layers.{b}", x, p, drop, s["ffn1"], s["ffn2"])
        x = x * keep
    return torch_layernorm(x, sd["sas.last_layernorm.weight"], sd["sas.last_layernorm.bias"], LN_EPS)


def forward(sd, seq, pos, neg, n_blocks, heads, **kw):
    """``SAS.forward`` NN/models/sas_model/sas.py:90-105 -> (pos_logits, neg_logits) [B, L]."""
    f = log2feats(sd, seq, n_blocks, heads, **kw)
    E = sd["sas.item_emb.weight"]
    return (f * F.embedding(pos, E, padding_idx=0)).sum(-1), (f * F.embedding(neg, E, padding_idx=0)).sum(-1)


def bce_pair_loss(pos_logits: torch.Tensor, neg_logits: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """NN/trainers/sas.py:38-49: BCEWithLogits(pos, 1) + BCEWithLogits(neg, 0), each a mean over the
    positions where ``pos != 0``."""
    idx = pos != 0
    pl, nl = pos_logits[idx], neg_logits[idx]
    return F.binary_cross_entropy_with_logits(pl, torch.ones_like(pl)) + \
        F.binary_cross_entropy_with_logits(nl, torch.zeros_like(nl))


def loss(sd, seq, pos, neg, n_blocks, heads, l2_emb: float = 0.0, **kw) -> torch.Tensor:
    """``SASTrainer.calculate_loss`` NN/trainers/sas.py:34-54 (incl. ``l2_emb * ||param||_2`` summed
    over ALL parameters, :51-52)."""
    pl, nl = forward(sd, seq, pos, neg, n_blocks, heads, **kw)
    out = bce_pair_loss(pl, nl, pos)
    if l2_emb != 0.0:
        for p_ in sd.values():
            out = out + l2_emb * torch.norm(p_)
    return out


def predict(sd, seq, item_indices, n_blocks, heads) -> torch.Tensor:
    """``SAS.predict`` NN/models/sas_model/sas.py:107-118 -> [B, C]."""
    f = log2feats(sd, seq, n_blocks, heads)[:, -1, :]
    return F.embedding(item_indices, sd["sas.item_emb.weight"], padding_idx=0).matmul(f.unsqueeze(-1)).squeeze(-1)


def scores_full_catalogue(sd, seq, n_blocks, heads, chunk: int = 1 << 20) -> torch.Tensor:
    """Full-catalogue restatement of :func:`predict` with candidates = all items 1..V, chunked over V
    (mathematically identical to gathering every row) -> [B, V] (column j = item j+1)."""
    f = log2feats(sd, seq, n_blocks, heads)[:, -1, :]
    E = sd["sas.item_emb.weight"]
    outs = [f @ E[v0:min(v0 + chunk, E.size(0))].t() for v0 in range(1, E.size(0), chunk)]
    return torch.cat(outs, dim=1)


def state_dict_shapes(num_items: int, max_len: int, d: int, n_blocks: int) -> Dict[str, tuple]:
    """Key names / shapes / ORDER of the reference ``SASModel.state_dict()`` (module registration order
    of NN/models/sas_model/sas.py:30-55; SURVEY.md 8b [probed])."""
    s = {"sas.item_emb.weight": (num_items + 1, d), "sas.pos_emb.weight": (max_len, d)}
    for b in range(n_blocks):
        s[f"sas.attention_layernorms.{b}.weight"] = (d,)
        s[f"sas.attention_layernorms.{b}.bias"] = (d,)
    for b in range(n_blocks):
        s[f"sas.attention_layers.{b}.in_proj_weight"] = (3 * d, d)
        s[f"sas.attention_layers.{b}.in_proj_bias"] = (3 * d,)
        s[f"sas.attention_layers.{b}.out_proj.weight"] = (d, d)
        s[f"sas.attention_layers.{b}.out_proj.bias"] = (d,)
    for b in range(n_blocks):
        s[f"sas.forward_layernorms.{b}.weight"] = (d,)
        s[f"sas.forward_layernorms.{b}.bias"] = (d,)
    for b in range(n_blocks):
        s[f"sas.forward_layers.{b}.conv1.weight"] = (d, d, 1)
        s[f"sas.forward_layers.{b}.conv1.bias"] = (d,)
        s[f"sas.forward_layers.{b}.conv2.weight"] = (d, d, 1)
        s[f"sas.forward_layers.{b}.conv2.bias"] = (d,)
    s["sas.last_layernorm.weight"] = (d,)
    s["sas.last_layernorm.bias"] = (d,)
    return s


def random_state_dict(num_items, max_len, d, n_blocks, seed: int = 0, scale: float = 0.1) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in state_dict_shapes(num_items, max_len, d, n_blocks).items():
        if "layernorm" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(*shp, generator=g)
        else:
            sd[k] = torch.randn(*shp, generator=g) * (1.0 if "emb" in k.split(".")[1] else scale)
    sd["sas.item_emb.weight"][0].zero_()
    return sd


def sample_batch(user_train, item_num: int, batch_size: int, max_len: int, rng: np.random.RandomState):
    """Wire format of ``sample_function`` NN/dataloaders/sas.py:70-86: (seq, pos, neg) int64 [B, L],
    left-padded with 0; one uniform negative per real position drawn from {0..V} \\ set(train)."""
    seqs, poss, negs = [], [], []
    for _ in range(batch_size):
        train = user_train[rng.randint(0, len(user_train))][-max_len:]
        pad = max_len - len(train) + 1
        allowed = np.array(sorted(set(range(0, item_num + 1)) - set(train)))
        neg = allowed[rng.randint(0, len(allowed), size=len(train) - 1)]
        seqs.append([0] * pad + list(train[:-1]))
        poss.append([0] * pad + list(train[1:]))
        negs.append([0] * pad + list(neg))
    return np.array(seqs, dtype=np.int64), np.array(poss, dtype=np.int64), np.array(negs, dtype=np.int64)
