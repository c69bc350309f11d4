"""Dense Adam / StepLR CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

The reference builds ``optim.Adam(model.parameters(), lr, weight_decay)`` with default betas
(0.9, 0.999), eps 1e-8 (NN/trainers/base.py:225-233) and steps it every batch (:123); tables are
DENSE parameters, so every row moves every step once its moments are non-zero.
``NN/`` = ``/root/reference/NerualNetwork/bert4rec&sas4rec/``.
"""
from __future__ import annotations

import math

import numpy as np


def adam_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, step: int, lr: float,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0):
    """One ``torch.optim.Adam`` step (amsgrad=False, maximize=False), fp32, in place; ``step`` is the
    1-based step count AFTER increment.  Follows torch's single-tensor formula:
    g += wd*p; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g;
    p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""
    f = np.float32
    g = g.astype(np.float32)
    if weight_decay != 0.0:
        g = g + f(weight_decay) * p
    m += (g - m) * f(1.0 - beta1)
    v *= f(beta2)
    v += f(1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
    p -= f(step_size) * (m / denom)
    return p, m, v


def step_lr(base_lr: float, epoch: int, decay_step: int, gamma: float) -> float:
    """``StepLR(step_size=decay_step, gamma)`` NN/trainers/base.py:40,87."""
    return base_lr * gamma ** (epoch // decay_step)
