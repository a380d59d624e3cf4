"""Oracle: optimizer arithmetic as the reference constructs it (pipeline.py:117-127, 241-244).

* ``torch.optim.AdamW(flat_params, lr, betas=(b1,b2), eps, fused=True)`` -- weight_decay is the
  torch default 0.01, amsgrad off, maximize off.  Restated from torch's documented update rule
  (third party; pinned against ``torch.optim.AdamW`` itself in tests/test_oracle_optim.py).
* deepspeed ``WarmupLR(optimizer, <get_warmup_steps()>)`` -- the reference passes the HF warm-up step
  count (0 by default) as the ctor's 2nd positional argument, which is ``warmup_min_lr``; every other
  argument keeps deepspeed's default (max 1e-3, 1000 steps, log).  deepspeed is not installable here:
  parity unpinned, restated from its published source.
"""
from __future__ import annotations

import math

import torch


class WarmupLR:
    def __init__(self, warmup_min_lr: float = 0.0, warmup_max_lr: float = 1e-3, warmup_num_steps: int = 1000):
        self.min_lr = float(warmup_min_lr)
        self.max_lr = float(warmup_max_lr)
        self.delta = self.max_lr - self.min_lr
        self.warmup_num_steps = max(2, warmup_num_steps)
        self.inverse_log_warm_up = 1.0 / math.log(self.warmup_num_steps)
        self.last_batch_iteration = -1

    def _gamma(self) -> float:
        if self.last_batch_iteration < self.warmup_num_steps:
            return self.inverse_log_warm_up * math.log(self.last_batch_iteration + 1)
        return 1.0

    def get_lr(self) -> float:
        if self.last_batch_iteration < 0:
            return self.min_lr  # deepspeed returns [0.0]; min_lr is 0 in the reference's call
        return self.min_lr + self.delta * self._gamma()

    def step(self) -> float:
        self.last_batch_iteration += 1
        return self.get_lr()


def lr_sequence(n: int, **kw) -> list[float]:
    """lr seen by the k-th ``optimizer.step()``: ctor leaves the optimizer at its own lr until the
    first ``scheduler.step()``; deepspeed's ctor calls ``step(last_batch_iteration)`` once, so the
    first optimizer step already sees ``get_lr()`` at iteration -1 -> 0."""
    s = WarmupLR(**kw)
    out = [s.get_lr()]
    for _ in range(n - 1):
        out.append(s.step())
    return out


def adamw_step_(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float,
                beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.01) -> None:
    """One AdamW update in place, fp32, ``step`` is 1-based."""
    p.mul_(1.0 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
