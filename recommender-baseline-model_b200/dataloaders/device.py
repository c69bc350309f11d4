"""Training-batch producers that run ON THE DEVICE.

The reference builds every batch with python per-item loops, ``deepcopy`` and (SASRec) ``multiprocessing.Queue`` pickling
(NN/dataloaders/bert.py:77-110, NN/dataloaders/sas.py:65-114).  Here the user histories live on the GPU as a CSR and a
batch is one kernel launch (``rbm_bert_cloze_batch`` / ``rbm_sas_train_batch``); what comes out has the reference's wire
format (int64 ``[B, L]`` tensors, left-padded with 0), so the trainers consume it unchanged.  A batch is a pure function
of (histories, users, seed, step): the Philox field layout is part of the C-ABI contract (include/rbm.h)."""
from typing import Dict, Sequence, Union

import numpy as np
import torch

from .. import ops


def histories_to_csr(user_train: Union[Dict[int, Sequence[int]], Sequence[Sequence[int]]], device):
    """``user_train`` of ``data_partition`` (dict user -> item list, or a list of lists) -> (hist_ptr, hist_items) on device."""
    rows = [user_train[u] for u in sorted(user_train)] if isinstance(user_train, dict) else list(user_train)
    ptr = np.zeros(len(rows) + 1, np.int64)
    np.cumsum([len(r) for r in rows], out=ptr[1:])
    items = np.fromiter((i for r in rows for i in r), np.int64, count=int(ptr[-1]))
    return torch.from_numpy(ptr).to(device), torch.from_numpy(items).to(device)


class DeviceBertTrainLoader:
    """Epoch iterator of ``(tokens, labels)`` like ``DataLoader(BertTrainDataset, shuffle=True)`` (NN/dataloaders/bert.py:25-34)."""

    def __init__(self, user_train, max_len: int, mask_prob: float, num_items: int, batch_size: int, device, seed: int = 0):
        self.ptr, self.items = histories_to_csr(user_train, device)
        self.num_users = self.ptr.numel() - 1
        self.max_len, self.mask_prob, self.num_items, self.batch_size = max_len, mask_prob, num_items, batch_size
        self.mask_token = num_items + 1
        self.seed, self.step = seed, 0
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)

    def __len__(self):
        return (self.num_users + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        perm = torch.randperm(self.num_users, device=self.ptr.device, generator=self.gen)
        for b in range(len(self)):
            users = perm[b * self.batch_size:(b + 1) * self.batch_size].contiguous()
            yield ops.bert_cloze_batch(self.ptr, self.items, users, self.max_len, self.mask_prob, self.mask_token, self.num_items,
                                       self.seed, self.step)
            self.step += 1


class DeviceSasTrainLoader:
    """``WarpSampler`` (NN/dataloaders/sas.py:92-121): ``len(user_train) // batch_size`` batches per epoch, every row a uniformly
    drawn user (with replacement), as ``(seq, pos, neg)`` device tensors."""

    def __init__(self, user_train, max_len: int, num_items: int, batch_size: int, device, seed: int = 0):
        self.ptr, self.items = histories_to_csr(user_train, device)
        self.num_users = self.ptr.numel() - 1
        self.max_len, self.num_items, self.batch_size = max_len, num_items, batch_size
        self.seed, self.step = seed, 0
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)

    def __len__(self):
        return self.num_users // self.batch_size

    def __iter__(self):
        for _ in range(len(self)):
            users = torch.randint(0, self.num_users, (self.batch_size,), device=self.ptr.device, generator=self.gen)
            yield ops.sas_train_batch(self.ptr, self.items, users, self.max_len, self.num_items, self.seed, self.step)
            self.step += 1
