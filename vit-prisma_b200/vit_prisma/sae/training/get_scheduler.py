"""Learning-rate schedules of the SAE trainer (reference sae/training/get_scheduler.py:17-92).

``get_scheduler`` keeps the reference signature and returns torch ``LambdaLR`` / cosine schedulers for callers
that own a torch optimizer.  The fused trainer needs only the scalar multiplier per step, exposed as
``lr_multiplier_fn`` -- the learning rate is a kernel argument, there is no optimizer object to mutate.

Quirk preserved: for ``cosineannealingwarmup`` the trainer passes ``lr_end = cfg.lr / 10`` and the lambda uses it as a
*multiplier* floor (get_scheduler.py:42-53, train_sae.py:235), so the final LR is ``lr * lr / 10``, not ``lr / 10``.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Optional

import torch.optim as optim
import torch.optim.lr_scheduler as lr_scheduler


def lr_multiplier_fn(scheduler_name: Optional[str], **kwargs: Any) -> Callable[[int], float]:
    name = (scheduler_name or "constant").lower()
    warm = kwargs.get("warm_up_steps", 0)
    total = kwargs.get("training_steps")
    lr_end = kwargs.get("lr_end", 0)
    if name == "constant":
        return lambda step: 1.0
    if name == "constantwithwarmup":
        return lambda step: min(1.0, (step + 1) / warm)
    if name == "linearwarmupdecay":
        assert total is not None, "training_steps must be provided"
        return lambda step: (step + 1) / warm if step < warm else (total - step) / (total - warm)
    if name == "cosineannealingwarmup":
        assert total is not None, "training_steps must be provided"

        def fn(step: int) -> float:
            if step < warm:
                return (step + 1) / warm
            progress = (step - warm) / (total - warm)
            return lr_end + 0.5 * (1 - lr_end) * (1 + math.cos(math.pi * progress))
        return fn
    raise ValueError(f"Unsupported scheduler for the fused trainer: {scheduler_name}")


def get_scheduler(scheduler_name: Optional[str], optimizer: optim.Optimizer, **kwargs: Any):
    name = (scheduler_name or "constant").lower()
    if name in ("constant", "constantwithwarmup", "linearwarmupdecay", "cosineannealingwarmup"):
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lr_multiplier_fn(name, **kwargs))
    if name == "cosineannealing":
        total = kwargs.get("training_steps")
        assert total is not None, "training_steps must be provided"
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=total, eta_min=kwargs.get("lr_end", 0))
    if name == "cosineannealingwarmrestarts":
        total = kwargs.get("training_steps")
        return lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=total // kwargs.get("num_cycles", 1), eta_min=kwargs.get("lr_end", 0))
    raise ValueError(f"Unsupported scheduler: {scheduler_name}")
