"""``data_partition`` (NN/dataloaders/__init__.py:17-66) without the python per-line loop.

The reference reads ``"<user> <item>"`` lines one by one into a ``defaultdict(list)`` -- minutes at 10^8 interactions.  Here the
file is parsed in one vectorised pass (pandas' C reader), users keep their order of first appearance and items their file order
(one stable sort), and the sliding-window / leave-one-out split is :func:`sliding_window_partition`; the result is the reference's
``[user_train, user_valid, user_test, usernum, itemnum]`` (pinned against the reference function in tests/golden/partition.npz).
"""
from typing import List

import numpy as np

from .synthetic import sliding_window_partition


def load_interactions_text(path: str) -> List[np.ndarray]:
    """Per-user item histories (users in order of first appearance, items in file order) and nothing else."""
    import pandas as pd
    df = pd.read_csv(path, sep=" ", header=None, names=["u", "i"], dtype=np.int64, engine="c")
    u, i = df["u"].to_numpy(), df["i"].to_numpy()
    if u.size == 0:
        return []
    uniq, first = np.unique(u, return_index=True)
    order = np.argsort(u, kind="stable")          # groups the lines of a user, file order kept inside a group
    counts = np.bincount(np.searchsorted(uniq, u), minlength=uniq.size)
    bounds = np.concatenate([[0], np.cumsum(counts)])
    items_sorted = i[order]
    by_first = np.argsort(first, kind="stable")   # dict insertion order of the reference = order of first appearance
    return [items_sorted[bounds[g]:bounds[g + 1]] for g in by_first]


def data_partition(path: str, max_len: int, prop_sliding_window: float):
    """Drop-in for ``data_partition(fname, max_len, prop_sliding_window)`` given the file's path: same five return values
    (``itemnum`` is the largest item id of the file, ``usernum`` the number of windows)."""
    hist = load_interactions_text(path)
    train, valid, test, n, _ = sliding_window_partition(hist, max_len, prop_sliding_window)
    itemnum = int(max((int(h.max()) for h in hist if len(h)), default=0))
    return [train, valid, test, n, itemnum]
