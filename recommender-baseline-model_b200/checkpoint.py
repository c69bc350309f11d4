"""Checkpoints in the reference's file format when the item tables are row-sharded over the ranks.

The reference writes ``{'model_state_dict', 'optimizer_state_dict', 'epoch'}`` with ``torch.save`` (NN/loggers.py:40-76,
NN/config.py:3-4, NN/trainers/base.py:255-259) and on resume loads the model part only -- the optimizer line is commented
out (NN/trainers/base.py:24-30).  Here:

* **gather on save** -- every rank hands in its row block ``[v_begin, v_end)`` of each sharded tensor (``ROW_SHARDED``) and
  of that tensor's Adam moments; rank 0 concatenates the blocks in rank order and writes ONE ``.pth`` whose tensors have
  the reference's full shapes, so the file loads into the reference (and into a single-GPU run) unchanged;
* **scatter on load** -- every rank opens the file memory-mapped (``torch.load(mmap=True)``: a 10 M x 64 table with its two
  Adam moments is 7.7 GB, no rank reads more than its block plus the small replicated tensors) and keeps its rows;
* **resume incl. optimizer state** -- ``exp_avg`` / ``exp_avg_sq`` / ``step`` travel with the parameters.

Everything here is host-side plumbing over ``torch.distributed`` (gloo or NCCL); no kernel is involved.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .dist import shard_range

STATE_DICT_KEY = "model_state_dict"            # NN/config.py:3
OPTIMIZER_STATE_DICT_KEY = "optimizer_state_dict"  # NN/config.py:4

# tensors whose dim 0 is the item axis (SURVEY.md 8b "checkpoint compat", 8e "row-sharded [v_begin, v_end)")
ROW_SHARDED = {
    "bert": ("bert.embedding.token.weight", "out.weight", "out.bias"),
    "sas": ("sas.item_emb.weight",),
}


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_rows(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's contiguous row block (``dist.shard_range`` over dim 0) as an owning tensor."""
    b, e = shard_range(t.shape[0], rank, world)
    return t[b:e].clone()


def shard_state_dict(full: Dict[str, torch.Tensor], keys: Iterable[str], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Full-shape model ``state_dict`` -> the same dict with the tensors named in ``keys`` cut to this rank's rows."""
    keys = set(keys)
    return {k: (shard_rows(v, rank, world) if k in keys else v) for k, v in full.items()}


def _gather_rows(local: torch.Tensor, total_rows: int, group=None) -> Optional[torch.Tensor]:
    """Row blocks of all ranks (rank order) -> full tensor on rank 0, None elsewhere.  Blocks are padded to the common
    ``ceil(total_rows / world)`` rows for the exchange."""
    rank, world = _world(group)
    if world == 1:
        return local
    per = (total_rows + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, parts, dst=0, group=group)
    if rank != 0:
        return None
    out = []
    for r, p in enumerate(parts):
        b, e = shard_range(total_rows, r, world)
        out.append(p[:e - b])
    return torch.cat(out)


def gather_state_dict(local: Dict[str, torch.Tensor], keys: Iterable[str], total_rows: Dict[str, int], group=None):
    """Inverse of :func:`shard_state_dict`: full-shape dict on rank 0 (CPU tensors), None on the other ranks."""
    rank, _ = _world(group)
    keys = set(keys)
    out = {}
    for k, v in local.items():  # same key order on every rank: the collectives line up
        if k in keys:
            g = _gather_rows(v.detach(), total_rows[k], group)
            if rank == 0:
                out[k] = g.cpu()
        elif rank == 0:
            out[k] = v.detach().cpu()
    return out if rank == 0 else None


def _param_names(model: torch.nn.Module, optimizer: torch.optim.Optimizer) -> List[str]:
    """Name of the parameter behind every index of ``optimizer.state_dict()['state']`` (torch numbers the parameters of all
    groups consecutively in group order)."""
    by_id = {id(p): n for n, p in model.named_parameters()}
    return [by_id[id(p)] for g in optimizer.param_groups for p in g["params"]]


def save_checkpoint(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer], epoch: int,
                    sharded_keys: Sequence[str] = (), total_rows: Optional[Dict[str, int]] = None, group=None,
                    extra_steps: int = 0) -> None:
    """Write the reference's checkpoint dict.  ``sharded_keys`` name the ``state_dict`` entries of which this rank's model
    holds only its row block; ``total_rows[key]`` is the full row count.  Collective: every rank calls it; rank 0 writes.
    ``extra_steps``: steps a live CUDA graph has taken on the device that the python-side counters have not seen
    (``trainer._graph_steps_done()``); folded into the saved Adam ``step`` and ``dropout_step``."""
    rank, world = _world(group)
    total_rows = dict(total_rows or {})
    msd = gather_state_dict(model.state_dict(), sharded_keys, total_rows, group)
    osd = None
    if optimizer is not None:
        names = _param_names(model, optimizer)
        raw = optimizer.state_dict()
        state = {}
        for idx in sorted(raw["state"]):
            st = raw["state"][idx]
            name = names[idx]
            ent = {}
            for f, v in st.items():
                if torch.is_tensor(v) and v.dim() > 0 and name in sharded_keys:
                    g = _gather_rows(v.detach(), total_rows[name], group)
                    ent[f] = g.cpu() if rank == 0 else None
                else:
                    ent[f] = v.detach().cpu() if torch.is_tensor(v) else v
                if f == "step" and extra_steps:
                    ent[f] = ent[f] + extra_steps
            state[idx] = ent
        osd = {"state": state, "param_groups": raw["param_groups"]}
    if rank == 0:
        tmp = path + ".tmp"
        torch.save({STATE_DICT_KEY: msd, OPTIMIZER_STATE_DICT_KEY: osd, "epoch": epoch,
                    "dropout_step": int(getattr(model, "_step", 0)) + extra_steps}, tmp)
        os.replace(tmp, path)  # a crash mid-write never leaves a truncated best_acc_model.pth behind
    if world > 1:
        dist.barrier(group=group)


def load_checkpoint(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
                    sharded_keys: Sequence[str] = (), group=None, map_location="cpu") -> int:
    """Load a reference-format checkpoint (written by the reference, by a single-GPU run or by :func:`save_checkpoint`) into
    a model that holds row blocks of ``sharded_keys``; with ``optimizer`` also the Adam state (the resume the reference leaves
    commented out).  Returns the stored epoch (-1 if absent)."""
    rank, world = _world(group)
    try:
        ck = torch.load(path, map_location="cpu", mmap=True, weights_only=False)
    except (RuntimeError, ValueError, TypeError):  # legacy (non-zip) files cannot be memory-mapped
        ck = torch.load(path, map_location="cpu", weights_only=False)
    keys = set(sharded_keys)
    msd = {k: (shard_rows(v, rank, world) if k in keys else v) for k, v in ck[STATE_DICT_KEY].items()}
    model.load_state_dict(msd)
    osd = ck.get(OPTIMIZER_STATE_DICT_KEY)
    if optimizer is not None and osd is not None:
        names = _param_names(model, optimizer)
        state = {}
        for idx, st in osd["state"].items():
            name = names[idx]
            state[idx] = {f: (shard_rows(v, rank, world) if (torch.is_tensor(v) and v.dim() > 0 and name in keys) else v)
                          for f, v in st.items()}
        optimizer.load_state_dict({"state": state, "param_groups": osd["param_groups"]})
    if "dropout_step" in ck and hasattr(model, "_step"):
        model._step = int(ck["dropout_step"])  # resume the dropout stream where it stopped (trainers/base.py _create_state_dict)
    return int(ck.get("epoch", -1))
