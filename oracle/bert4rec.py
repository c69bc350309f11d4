"""BERT4Rec CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Functional fp32 restatement over a reference-shaped ``state_dict`` (keys as produced by the
reference's ``BERTModel``: ``bert.embedding.token.weight`` ... ``out.bias``).
``NN/`` = ``/root/reference/NerualNetwork/bert4rec&sas4rec/``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .common import DropoutPlan, NO_DROPOUT, bert_layernorm, gelu_tanh

# dropout site numbering (shared convention with the CUDA path; see DESIGN.md "dropout")
SITE_EMB = 0


def block_sites(b: int):
    base = 1 + 5 * b
    return dict(attn=base, sub_in=base + 1, ffn=base + 2, sub_out=base + 3, block=base + 4)


def num_sites(n_blocks: int) -> int:
    return 1 + 5 * n_blocks


def attention(sd: Dict[str, torch.Tensor], pfx: str, x: torch.Tensor, key_mask: torch.Tensor, heads: int,
              p_attn: float, drop: DropoutPlan, site: int) -> torch.Tensor:
    """NN/models/bert_modules/attention/multi_head.py:24-40 + attention/single.py:13-35.

    ``key_mask`` [B, L] bool, True where token > 0 (NN/models/bert_modules/bert.py:38 broadcasts
    it over query rows and heads); masked keys get ``-1e9`` (not -inf) before the softmax.
    """
    B, L, d = x.shape
    dk = d // heads
    q, k, v = [F.linear(x, sd[f"{pfx}.linear_layers.{i}.weight"], sd[f"{pfx}.linear_layers.{i}.bias"])
               .view(B, L, heads, dk).transpose(1, 2) for i in range(3)]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    scores = scores.masked_fill(~key_mask[:, None, None, :], -1e9)
    p = F.softmax(scores, dim=-1)
    p = drop(p, p_attn, site)
    ctx = torch.matmul(p, v).transpose(1, 2).contiguous().view(B, L, d)
    return F.linear(ctx, sd[f"{pfx}.output_linear.weight"], sd[f"{pfx}.output_linear.bias"])


def feed_forward(sd, pfx: str, x: torch.Tensor, p_hidden: float, drop: DropoutPlan, site: int) -> torch.Tensor:
    """NN/models/bert_modules/utils/feed_forward.py:15-16: w_2(dropout(gelu(w_1 x)))."""
    u = F.linear(x, sd[f"{pfx}.w_1.weight"], sd[f"{pfx}.w_1.bias"])
    u = drop(gelu_tanh(u), p_hidden, site)
    return F.linear(u, sd[f"{pfx}.w_2.weight"], sd[f"{pfx}.w_2.bias"])


def hidden_states(sd: Dict[str, torch.Tensor], tokens: torch.Tensor, n_blocks: int, heads: int,
                  p_attn: float = 0.0, p_hidden: float = 0.0, drop: DropoutPlan = NO_DROPOUT) -> torch.Tensor:
    """``BERT.forward`` NN/models/bert_modules/bert.py:36-43 -> [B, L, d].

    embedding (NN/models/bert_modules/embedding/bert.py:29-31, position.py:14-16): token row +
    the whole ``pe.weight`` (so L must equal max_len), no sqrt(d) scale, no pad zeroing; then per
    block (transformer.py:28-32, sublayer.py:16-18): x += drop(attn(LN(x))); x += drop(ffn(LN(x)));
    x = drop(x).
    """
    key_mask = tokens > 0
    x = F.embedding(tokens, sd["bert.embedding.token.weight"], padding_idx=0) + sd["bert.embedding.position.pe.weight"].unsqueeze(0)
    x = drop(x, p_hidden, SITE_EMB)
    for b in range(n_blocks):
        s = block_sites(b)
        pfx = f"bert.transformer_blocks.{b}"
        n1 = bert_layernorm(x, sd[f"{pfx}.input_sublayer.norm.a_2"], sd[f"{pfx}.input_sublayer.norm.b_2"])
        a = attention(sd, f"{pfx}.attention", n1, key_mask, heads, p_attn, drop, s["attn"])
        x = x + drop(a, p_hidden, s["sub_in"])
        n2 = bert_layernorm(x, sd[f"{pfx}.output_sublayer.norm.a_2"], sd[f"{pfx}.output_sublayer.norm.b_2"])
        f = feed_forward(sd, f"{pfx}.feed_forward", n2, p_hidden, drop, s["ffn"])
        x = x + drop(f, p_hidden, s["sub_out"])
        x = drop(x, p_hidden, s["block"])
    return x


def logits(sd, tokens, n_blocks, heads, **kw) -> torch.Tensor:
    """``BERTModel.forward`` NN/models/bert.py:15-16: untied ``out`` Linear(d -> V+1) -> [B, L, V+1]."""
    h = hidden_states(sd, tokens, n_blocks, heads, **kw)
    return F.linear(h, sd["out.weight"], sd["out.bias"])


def loss(sd, tokens: torch.Tensor, labels: torch.Tensor, n_blocks: int, heads: int, **kw) -> torch.Tensor:
    """``BERTTrainer.calculate_loss`` NN/trainers/bert.py:30-41: CE(ignore_index=0) over all
    B*L rows of the flattened logits, mean over labels != 0."""
    lg = logits(sd, tokens, n_blocks, heads, **kw)
    return F.cross_entropy(lg.view(-1, lg.size(-1)), labels.view(-1), ignore_index=0)


def loss_masked_only(sd, tokens, labels, n_blocks, heads, chunk: int = 65536, **kw) -> torch.Tensor:
    """Mathematically identical chunked restatement of :func:`loss` that scores only the rows with
    labels != 0 (ignored rows contribute exactly 0 to loss and gradients, NN/trainers/bert.py:11).
    This is the form the CUDA path computes and the one usable at cfg4 sizes."""
    h = hidden_states(sd, tokens, n_blocks, heads, **kw)
    rows = labels.view(-1).nonzero(as_tuple=True)[0]
    hc = h.view(-1, h.size(-1))[rows]
    tgt = labels.view(-1)[rows]
    W, bias = sd["out.weight"], sd["out.bias"]
    m = torch.full((hc.size(0),), -float("inf"))
    s = torch.zeros(hc.size(0))
    for v0 in range(0, W.size(0), chunk):
        lg = F.linear(hc, W[v0:v0 + chunk], bias[v0:v0 + chunk])
        m_new = torch.maximum(m, lg.max(dim=1).values)
        s = s * torch.exp(m - m_new) + torch.exp(lg - m_new[:, None]).sum(1)
        m = m_new
    lse = m + torch.log(s)
    tl = (hc * W[tgt]).sum(1) + bias[tgt]
    return (lse - tl).mean()


def scores_last(sd, tokens, n_blocks, heads) -> torch.Tensor:
    """Eval scoring of ``BERTTrainer.calculate_metrics`` NN/trainers/bert.py:43-49 before the
    candidate gather: logits of the LAST position only -> [B, V+1] (the reference computes all L
    positions and keeps one)."""
    h = hidden_states(sd, tokens, n_blocks, heads)[:, -1, :]
    return F.linear(h, sd["out.weight"], sd["out.bias"])


def candidate_scores(sd, tokens, candidates, n_blocks, heads) -> torch.Tensor:
    """NN/trainers/bert.py:47-49: ``scores[:, -1, :].gather(1, candidates)``."""
    return scores_last(sd, tokens, n_blocks, heads).gather(1, candidates)


def state_dict_shapes(num_items: int, max_len: int, d: int, n_blocks: int) -> Dict[str, tuple]:
    """Key names / shapes of the reference ``BERTModel.state_dict()`` (SURVEY.md 8b [probed])."""
    s = {"bert.embedding.token.weight": (num_items + 2, d), "bert.embedding.position.pe.weight": (max_len, d)}
    for b in range(n_blocks):
        p = f"bert.transformer_blocks.{b}"
        for i in range(3):
            s[f"{p}.attention.linear_layers.{i}.weight"] = (d, d)
            s[f"{p}.attention.linear_layers.{i}.bias"] = (d,)
        s[f"{p}.attention.output_linear.weight"] = (d, d)
        s[f"{p}.attention.output_linear.bias"] = (d,)
        s[f"{p}.feed_forward.w_1.weight"] = (4 * d, d)
        s[f"{p}.feed_forward.w_1.bias"] = (4 * d,)
        s[f"{p}.feed_forward.w_2.weight"] = (d, 4 * d)
        s[f"{p}.feed_forward.w_2.bias"] = (d,)
        for n in ("input_sublayer", "output_sublayer"):
            s[f"{p}.{n}.norm.a_2"] = (d,)
            s[f"{p}.{n}.norm.b_2"] = (d,)
    s["out.weight"] = (num_items + 1, d)
    s["out.bias"] = (num_items + 1,)
    return s


def random_state_dict(num_items, max_len, d, n_blocks, seed: int = 0, scale: float = 0.1) -> Dict[str, torch.Tensor]:
    """Random-init weights of the reference architecture (test/bench input, not the reference's init)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in state_dict_shapes(num_items, max_len, d, n_blocks).items():
        if k.endswith("a_2"):
            sd[k] = 1.0 + 0.1 * torch.randn(*shp, generator=g)
        else:
            sd[k] = torch.randn(*shp, generator=g) * (1.0 if "embedding" in k else scale)
    sd["bert.embedding.token.weight"][0].zero_()
    return sd
