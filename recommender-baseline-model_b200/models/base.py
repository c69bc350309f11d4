"""BaseModel -- same contract as NN/models/base.py:6-14, plus the dropout-stream bookkeeping shared by both models."""
from abc import ABCMeta, abstractmethod

import torch
import torch.nn as nn

SITES_PER_STEP = 64  # dropout sites are numbered step*SITES_PER_STEP + local site (DESIGN.md "dropout")


class BaseModel(nn.Module, metaclass=ABCMeta):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.dropout_seed = int(getattr(args, "dropout_seed", None) or torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
        self._step = 0

    @classmethod
    @abstractmethod
    def code(cls):
        pass

    def _next_site_base(self) -> int:
        """Every training-mode forward gets a fresh block of dropout sites; the backward regenerates the same masks."""
        base = self._step * SITES_PER_STEP
        self._step += 1
        return base

    def _device_long(self, x) -> torch.Tensor:
        """numpy / CPU tensors in, int64 tensor on the parameters' device out (NN/models/sas_model/sas.py:60,93-94,112)."""
        dev = next(self.parameters()).device
        if not torch.is_tensor(x):
            x = torch.as_tensor(x)
        return x.to(device=dev, dtype=torch.long, non_blocking=True).contiguous()
