"""FusedAdam -- ``torch.optim.Adam`` semantics (NN/trainers/base.py:225-233: default betas, eps 1e-8, dense update of
EVERY parameter incl. whole item tables) as ONE multi-tensor kernel launch per step (rbm_adam_multi).

State layout (``state[p] = {step, exp_avg, exp_avg_sq}``) and ``param_groups`` match torch's Adam, so ``StepLR``
(NN/trainers/base.py:40), ``get_lr`` (:95-97) and checkpoint dicts (:255-259) work unchanged."""
from __future__ import annotations

import torch

from . import lib as L
from .lib import check, ptr, stream, count_launches


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._table_cache = {}

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)

    @staticmethod
    def _chunk_map(sizes, device):
        rows = []
        for ti, n in enumerate(sizes):
            for ci in range((n + L.ADAM_CHUNK - 1) // L.ADAM_CHUNK):
                rows.append((ti, ci))
        return torch.tensor(rows, dtype=torch.int32).reshape(-1, 2).to(device)

    @torch.no_grad()
    def init_state(self):
        """Create the (step, exp_avg, exp_avg_sq) state of every parameter now (instead of lazily inside the first step):
        needed before CUDA-graph capture, where a lazily created zero buffer would be re-zeroed by every replay."""
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)

    def _tables(self, gi, group, advance: bool):
        """(descriptor table, chunk map, step) of one parameter group; device copies are cached while the addresses of
        parameters, gradients and state stay the same (the caching allocator hands the same gradient blocks back step after
        step: no host->device copy per step, and none at all inside a captured CUDA graph once ``stage_tables`` has run)."""
        plist = [p for p in group["params"] if p.grad is not None]
        if not plist:
            return None
        L.require_cuda(*plist)
        dev = plist[0].device
        steps = set()
        table = []
        for p in plist:
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                raise RuntimeError("FusedAdam handles dense fp32 parameters only")
            if not p.is_contiguous():
                raise RuntimeError("FusedAdam needs contiguous parameters")
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if advance:
                st["step"] += 1
            steps.add(int(st["step"]))
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            if g is not p.grad:
                p.grad = g
            table.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()))
        if len(steps) != 1:
            raise RuntimeError("FusedAdam: parameters of one group must share their step count")
        sizes = tuple(t[4] for t in table)
        key = (gi, sizes, str(dev))
        cmap = self._table_cache.get(key)
        if cmap is None:
            cmap = self._chunk_map(sizes, dev)
            self._table_cache = {key: cmap}
        cache = self.__dict__.setdefault("_desc_cache", {})
        dkey = tuple(table)  # rbm_adam_tensor[] : 5 x 8-byte fields
        if gi not in cache or cache[gi][0] != dkey:
            if torch.cuda.is_current_stream_capturing():
                # gradients were allocated inside the capture (their addresses are final): stage the table through the pinned
                # buffer of prepare_capture() -- a host->device copy node that every replay repeats (1.4 KB)
                pinned, devbuf = self._capture_bufs[gi]
                n = len(table)
                pinned[:n] = torch.tensor(table, dtype=torch.int64)
                devbuf[:n].copy_(pinned[:n], non_blocking=True)
                cache[gi] = (dkey, devbuf[:n])
            else:
                cache[gi] = (dkey, torch.tensor(table, dtype=torch.int64).to(dev))
        return cache[gi][1], cmap, steps.pop()

    def prepare_capture(self):
        """Pinned host + device buffers for the descriptor tables, allocated BEFORE a CUDA-graph capture (no allocation of
        pinned memory while capturing); see ``_tables``."""
        self._capture_bufs = {}
        for gi, group in enumerate(self.param_groups):
            n = max(1, len(group["params"]))
            dev = group["params"][0].device
            self._capture_bufs[gi] = (torch.empty(n, 5, dtype=torch.int64).pin_memory(), torch.empty(n, 5, dtype=torch.int64, device=dev))
        self.__dict__.setdefault("_desc_cache", {}).clear()

    @torch.no_grad()
    def stage_tables(self):
        """Upload the tables for the CURRENT gradient buffers without taking a step (before CUDA-graph capture)."""
        for gi, group in enumerate(self.param_groups):
            self._tables(gi, group, advance=False)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = L.load()
        for gi, group in enumerate(self.param_groups):
            t = self._tables(gi, group, advance=True)
            if t is None:
                continue
            desc, cmap, step = t
            b1, b2 = group["betas"]
            check(lib.rbm_adam_multi(ptr(desc), ptr(cmap), cmap.shape[0], float(group["lr"]), float(b1), float(b2),
                                     float(group["eps"]), float(group["weight_decay"]), step, stream()), "adam_multi")
            count_launches()
        return loss
