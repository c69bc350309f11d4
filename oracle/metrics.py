"""Ranking / top-k CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

``NN/`` = ``/root/reference/NerualNetwork/bert4rec&sas4rec/``.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def canonical_rank(scores: torch.Tensor) -> torch.Tensor:
    """The canonical tie-break the whole build uses: score descending, candidate index ascending.

    The reference ranks with ``(-scores).argsort(dim=1)`` (NN/trainers/utils.py:36), which is not
    stable on ties; ``argsort(stable=True)`` of ``-scores`` is the canonicalised form (SURVEY.md 7.2
    hard part 1).  Without ties in the top-(k+1) both agree exactly."""
    return (-scores).argsort(dim=1, stable=True)


def recalls_ndcgs_and_mrr_for_ks(scores: torch.Tensor, labels: torch.Tensor, ks: Iterable[int],
                                 per_user: bool = False) -> Dict[str, float]:
    """NN/trainers/utils.py:28-57, with the canonical tie-break.  ``scores`` [B,C] f32, ``labels``
    [B,C] i64 (1 = relevant).  Returns batch means as python floats (or per-user tensors)."""
    metrics = {}
    scores = scores.cpu()
    labels = labels.cpu()
    answer_count = labels.sum(1)
    answer_count_float = answer_count.float()
    labels_float = labels.float()
    cut = canonical_rank(scores)
    for k in sorted(ks, reverse=True):
        cut = cut[:, :k]
        hits = labels_float.gather(1, cut)
        recall = hits.sum(1) / answer_count_float
        weights = 1 / torch.log2(torch.arange(2, 2 + k).float())
        dcg = (hits * weights).sum(1)
        idcg = torch.Tensor([weights[:min(int(n), k)].sum() for n in answer_count])
        ndcg = dcg / idcg
        mrr = (hits * (1 / torch.arange(1, k + 1).float())).sum(1)
        if per_user:
            metrics["Recall@%d" % k], metrics["NDCG@%d" % k], metrics["MRR@%d" % k] = recall, ndcg, mrr
        else:
            metrics["Recall@%d" % k] = recall.mean().item()
            metrics["NDCG@%d" % k] = ndcg.mean().item()
            metrics["MRR@%d" % k] = mrr.mean().item()
    return metrics


def scores_fp32_sequential(f: torch.Tensor, table: torch.Tensor, bias=None) -> torch.Tensor:
    """The fp32 score the full-catalogue ranking is defined on: ``s = 0; for c in range(d): s = fma(f[u,c], table[v,c], s)``
    in fp32, then ``+ bias[v]`` -- the reference's ``matmul`` (NN/models/sas_model/sas.py:114, NN/models/bert.py:16) with
    the summation order pinned, so that "top-k of the fp32 scores" is one well-defined answer.  The FMA is emulated in
    fp64 (the product of two fp32 values is exact there) and rounded to fp32 after every step.  [U,d] x [V,d] -> [U,V];
    runs on whatever device the tensors live on."""
    f64, t64 = f.double(), table.double()
    acc = torch.zeros(f.shape[0], table.shape[0], dtype=torch.float32, device=f.device)
    for c in range(f.shape[1]):
        acc = (acc.double() + f64[:, c:c + 1] * t64[:, c].unsqueeze(0)).float()
    if bias is not None:
        acc = acc + bias.float().unsqueeze(0)
    return acc


def topk_canonical(scores: np.ndarray, k: int, id_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k per row of ``scores`` [U,C] under (score desc, index asc); returns (values f32 [U,k],
    ids i64 [U,k]) with ids = column + id_offset.  Rows shorter than k are padded (-inf, -1)."""
    scores = np.asarray(scores, dtype=np.float32)
    U, C = scores.shape
    kk = min(k, C)
    order = np.argsort(-scores, axis=1, kind="stable")[:, :kk]
    vals = np.take_along_axis(scores, order, axis=1)
    ids = order.astype(np.int64) + id_offset
    if kk < k:
        vals = np.concatenate([vals, np.full((U, k - kk), -np.inf, np.float32)], 1)
        ids = np.concatenate([ids, np.full((U, k - kk), -1, np.int64)], 1)
    return vals, ids


def topk_merge(vals: np.ndarray, ids: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge per-shard top-k lists ``vals``/``ids`` [S,U,k] (global ids) into the global top-k under
    the same rule (score desc, id asc); padded entries (id < 0) sort last.  Shard-count invariant."""
    S, U, kk = vals.shape
    v = np.transpose(vals, (1, 0, 2)).reshape(U, S * kk)
    i = np.transpose(ids, (1, 0, 2)).reshape(U, S * kk)
    out_v = np.empty((U, k), np.float32)
    out_i = np.empty((U, k), np.int64)
    for u in range(U):
        big = np.where(i[u] < 0, np.iinfo(np.int64).max, i[u])
        order = np.lexsort((big, -v[u].astype(np.float64)))[:k]
        out_v[u], out_i[u] = v[u][order], i[u][order]
    return out_v, out_i


def full_catalogue_metrics(top_ids: np.ndarray, positives: np.ndarray, ks: Iterable[int]) -> Dict[str, np.ndarray]:
    """Per-user HR/NDCG/MRR@k from a top-K id list and ONE held-out positive per user -- exactly what
    NN/trainers/utils.py:41-55 reduces to when labels are one-hot: Recall@k = hit, NDCG@k =
    1/log2(rank+2) (idcg = 1), MRR@k = 1/(rank+1).  Returns per-user fp32 arrays."""
    top_ids = np.asarray(top_ids)
    U, K = top_ids.shape
    hit = top_ids == np.asarray(positives).reshape(U, 1)
    w_ndcg = (1 / torch.log2(torch.arange(2, 2 + K).float())).numpy()
    w_mrr = (1 / torch.arange(1, K + 1).float()).numpy()
    out = {}
    for k in ks:
        h = hit[:, :k]
        out["Recall@%d" % k] = h.any(1).astype(np.float32)
        out["NDCG@%d" % k] = (h * w_ndcg[:k]).sum(1).astype(np.float32)
        out["MRR@%d" % k] = (h * w_mrr[:k]).sum(1).astype(np.float32)
    return out
