"""One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch): data-parallel gradient all-reduce through one
flat bucket, and the vocab-shard / top-k exchange helpers.  The reference has no counterpart (only nn.DataParallel,
NN/trainers/base.py:32-34); the contract is "N ranks == 1 rank on the same global batch" (SURVEY.md 8e)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

from . import lib as L
from .lib import check, ptr, stream, count_launches


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block [begin, end) of a table row-sharded over `world` ranks (ceil(n/world) rows each)."""
    per = (n_rows + world - 1) // world
    b = min(rank * per, n_rows)
    return b, min(b + per, n_rows)


def split_batch(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows of a global batch owned by `rank` (strong scaling: the global batch is fixed)."""
    return shard_range(n, rank, world)


def bucket_layout(sizes: Sequence[int], align: int = 4) -> Tuple[List[int], int]:
    """Element offsets of each tensor inside the flat bucket (each start aligned to `align` elements) and total size."""
    offs, o = [], 0
    for n in sizes:
        offs.append(o)
        o += (n + align - 1) // align * align
    return offs, o


def chunk_map(sizes: Sequence[int]) -> torch.Tensor:
    rows = [(ti, ci) for ti, n in enumerate(sizes) for ci in range((n + L.ADAM_CHUNK - 1) // L.ADAM_CHUNK)]
    return torch.tensor(rows, dtype=torch.int32).reshape(-1, 2)


class GradSync:
    """Averages the gradients of `params` over all ranks: pack -> ONE all-reduce of a flat fp32 bucket -> unpack."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.sizes = [p.numel() for p in self.params]
        self.offsets, self.total = bucket_layout(self.sizes)
        dev = self.params[0].device
        self.bucket = torch.zeros(self.total, device=dev, dtype=torch.float32)
        self.cmap = chunk_map(self.sizes).to(dev)

    def allreduce_bucket(self):
        if self.world > 1:
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.group)

    def _table(self):
        rows = []
        for p, n, o in zip(self.params, self.sizes, self.offsets):
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            p.grad = g
            rows.append((g.data_ptr(), n, o))
        # gradient buffers keep their addresses from step to step: the device copy of the table is reused (a pageable
        # host->device copy every step would block the host until the stream has drained)
        key = tuple(r[0] for r in rows)
        if getattr(self, "_tab_key", None) != key:
            self._tab = torch.tensor(rows, dtype=torch.int64).to(self.bucket.device)
            self._tab_key = key
        return self._tab

    def allreduce_grads(self):
        if self.world == 1:
            return
        lib = L.load()
        tab = self._table()
        n = self.cmap.shape[0]
        check(lib.rbm_bucket_pack(ptr(tab), ptr(self.cmap), n, ptr(self.bucket), 1.0 / self.world, 0, stream()), "bucket_pack")
        self.allreduce_bucket()
        check(lib.rbm_bucket_pack(ptr(tab), ptr(self.cmap), n, ptr(self.bucket), 1.0, 1, stream()), "bucket_unpack")
        count_launches(2)


def gather_topk(vals: torch.Tensor, ids: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-shard top-k lists [U,k] -> [S,U,k] (80 B per user per shard for k=10)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return vals.unsqueeze(0), ids.unsqueeze(0)
    vs = [torch.empty_like(vals) for _ in range(world)]
    is_ = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(vs, vals.contiguous(), group=group)
    dist.all_gather(is_, ids.contiguous(), group=group)
    return torch.stack(vs), torch.stack(is_)


def sharded_full_catalogue_topk(last_hidden: torch.Tensor, table: torch.Tensor, bias, num_items: int, k: int, group=None):
    """Vocab-parallel top-k: every rank scores ALL users against ITS row block of the (replicated or locally held)
    table with global item ids, the per-shard lists are exchanged and merged with the same (score desc, id asc) rule,
    so the result does not depend on the number of shards."""
    from . import ops
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b, e = shard_range(num_items, rank, world)  # items are rows 1..V
    if e > b:
        vals, ids = ops.score_topk(last_hidden, table, bias, 1 + b, 1 + e, k)
    else:
        U = last_hidden.shape[0]
        vals = torch.full((U, k), float("-inf"), device=last_hidden.device)
        ids = torch.full((U, k), -1, device=last_hidden.device, dtype=torch.int64)
    gv, gi = gather_topk(vals, ids, group)
    return ops.topk_merge(gv, gi)
