"""Synthetic interaction data and vectorised batch builders.

Wire formats (identical to the reference's datasets):
  BERT train  (tokens i64 [B,L], labels i64 [B,L]) left-padded with 0, [MASK] = V+1, label = original item at
              masked positions else 0; per item: mask w.p. ``mask_prob``; of masked 80% -> [MASK], 10% -> uniform
              random item, 10% keep          (BertTrainDataset.__getitem__ NN/dataloaders/bert.py:77-110)
  BERT eval   (seq i64 [B,L] ending in [MASK], candidates i64 [B,1+N] positive first, labels i64 [B,1+N])
                                              (BertEvalDataset.__getitem__ NN/dataloaders/bert.py:128-142)
  SAS train   (seq, pos, neg) i64 [B,L]: seq = train[:-1], pos = train[1:], one uniform negative per real position
              from {0..V} \\ set(train)       (sample_function NN/dataloaders/sas.py:70-86)
  SAS eval    (seq [B,L], candidates [B,1+N], labels [B,1+N])   (SASEvalDataset NN/dataloaders/sas.py:136-153)
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def _zipf_sampler(num_items: int, alpha: float):
    w = 1.0 / np.power(np.arange(1, num_items + 1, dtype=np.float64), alpha)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]

    def draw(rng: np.random.RandomState, n: int) -> np.ndarray:
        return (np.searchsorted(cdf, rng.rand(n), side="left") + 1).astype(np.int64)

    return draw


def synthetic_interactions(num_users: int, num_items: int, mean_len: float, min_len: int, max_user_len: int = 2000,
                           zipf_alpha: float = 1.0, seed: int = 1234, uniform_len: int = 0) -> List[np.ndarray]:
    """Per-user item-id histories (1-based ids, Zipf(alpha) popularity, no immediate repeats, clipped log-normal lengths
    -- or exactly ``uniform_len`` when given)."""
    rng = np.random.RandomState(seed)
    draw = _zipf_sampler(num_items, zipf_alpha)
    if uniform_len:
        lens = np.full(num_users, uniform_len, np.int64)
    else:
        sigma = 0.9
        mu = np.log(max(mean_len, 1.0)) - 0.5 * sigma * sigma
        lens = np.clip(np.round(rng.lognormal(mu, sigma, size=num_users)), min_len, max_user_len).astype(np.int64)
    flat = draw(rng, int(lens.sum()))
    while True:  # break immediate repeats (a replacement can collide with its right neighbour: iterate)
        rep = np.flatnonzero(flat[1:] == flat[:-1]) + 1
        if rep.size == 0:
            break
        flat[rep] = (flat[rep] + rng.randint(0, num_items - 1, size=rep.size)) % num_items + 1
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + n])
        o += n
    return out


def sliding_window_partition(histories: Sequence[np.ndarray], max_len: int, prop_sliding_window: float = 0.3):
    """The windowing + leave-one-out split of ``data_partition`` (NN/dataloaders/__init__.py:38-66) applied to in-memory
    histories: sequences longer than max_len are cut into overlapping windows (step int(prop*max_len)), each window an
    independent "user"; train = window[:-2], valid = [window[-2]], test = [window[-1]]."""
    step = int(prop_sliding_window * max_len) if prop_sliding_window != -1.0 else max_len
    data = []
    for h in histories:
        n = len(h)
        if n < 3:
            continue
        if n <= max_len:
            data.append(h)
        else:
            for i in list(range(n - max_len, 0, -step))[::-1]:
                data.append(h[i:i + max_len])
    train = [list(map(int, w[:-2])) for w in data]
    valid = [[int(w[-2])] for w in data]
    test = [[int(w[-1])] for w in data]
    return [train, valid, test, len(data), int(max(int(np.max(h)) for h in histories))]


def _left_pad(rows: Sequence[Sequence[int]], max_len: int) -> np.ndarray:
    out = np.zeros((len(rows), max_len), np.int64)
    for b, r in enumerate(rows):
        r = r[-max_len:]
        if len(r):
            out[b, max_len - len(r):] = r
    return out


class BertBatcher:
    """Vectorised Cloze-masking batches (same per-item rule as NN/dataloaders/bert.py:84-99)."""

    def __init__(self, user_train, num_items: int, max_len: int, mask_prob: float, seed: int = 0):
        self.seqs = _left_pad(user_train, max_len)
        self.num_items, self.max_len, self.mask_prob = num_items, max_len, mask_prob
        self.mask_token = num_items + 1
        self.rng = np.random.RandomState(seed)

    def __len__(self):
        return self.seqs.shape[0]

    def batch(self, batch_size: int) -> Tuple[np.ndarray, np.ndarray]:
        idx = self.rng.randint(0, len(self), size=batch_size)
        seq = self.seqs[idx]
        real = seq != 0
        prob = self.rng.rand(*seq.shape)
        masked = real & (prob < self.mask_prob)
        sub = prob / self.mask_prob
        tokens = seq.copy()
        tokens[masked & (sub < 0.8)] = self.mask_token
        rnd = masked & (sub >= 0.8) & (sub < 0.9)
        tokens[rnd] = self.rng.randint(1, self.num_items + 1, size=int(rnd.sum()))
        labels = np.where(masked, seq, 0)
        return tokens, labels


class SasBatcher:
    """(seq, pos, neg) batches in the format of ``sample_function`` (NN/dataloaders/sas.py:70-86)."""

    def __init__(self, user_train, num_items: int, max_len: int, seed: int = 0):
        self.user_train = [np.asarray(u, dtype=np.int64) for u in user_train if len(u) >= 2]
        self.num_items, self.max_len = num_items, max_len
        self.rng = np.random.RandomState(seed)

    def __len__(self):
        return len(self.user_train)

    def batch(self, batch_size: int):
        L = self.max_len
        seq = np.zeros((batch_size, L), np.int64)
        pos = np.zeros((batch_size, L), np.int64)
        neg = np.zeros((batch_size, L), np.int64)
        for b in range(batch_size):
            train = self.user_train[self.rng.randint(0, len(self.user_train))][-(L + 1):]
            n = len(train) - 1
            seq[b, L - n:] = train[:-1]
            pos[b, L - n:] = train[1:]
            # uniform over {0..V} \ set(train) by rejection (same distribution as random_neq, NN/dataloaders/sas.py:65-67)
            cand = self.rng.randint(0, self.num_items + 1, size=n)
            seen = np.isin(cand, train)
            while seen.any():
                cand[seen] = self.rng.randint(0, self.num_items + 1, size=int(seen.sum()))
                seen = np.isin(cand, train)
            neg[b, L - n:] = cand
        return seq, pos, neg


def eval_sequences(user_train, answers_prev, max_len: int, mask_token: int = 0) -> np.ndarray:
    """Evaluation inputs [U, L]: history (+ the validation item when ``answers_prev`` is given, i.e. test mode) and,
    for BERT, a trailing [MASK] (NN/dataloaders/bert.py:50-57,137-140; NN/dataloaders/sas.py:52-59,148-151)."""
    rows = []
    for u, seq in enumerate(user_train):
        s = list(seq)
        if answers_prev is not None:
            s = s + [answers_prev[u][0]]
        if mask_token:
            s = s + [mask_token]
        rows.append(s)
    return _left_pad(rows, max_len)


def uniform_negative_candidates(answers, num_items: int, n_neg: int, seed: int = 98765):
    """candidates = [answer] + n_neg uniform negatives != answer, labels = [1, 0, ...] (the reference's 'random'
    negative sampler protocol, NN/dataloaders/negative_samplers/random.py:17-37, without the seen-item exclusion)."""
    rng = np.random.RandomState(seed)
    ans = np.asarray([a[0] for a in answers], np.int64)
    negs = rng.randint(1, num_items, size=(len(ans), n_neg)).astype(np.int64)
    negs += negs >= ans[:, None]
    cands = np.concatenate([ans[:, None], negs], 1)
    labels = np.zeros_like(cands)
    labels[:, 0] = 1
    return cands, labels
