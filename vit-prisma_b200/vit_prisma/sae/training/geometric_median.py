"""Weiszfeld geometric median for the ``b_dec`` initialisation (reference sae/training/geometric_median.py:23-85).
Runs once before training on the first activation buffer; plain tensor ops on whatever device the points live on."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch


def weighted_average(points: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    weights = weights / weights.sum()
    return (points * weights.view(-1, 1)).sum(dim=0)


@torch.no_grad()
def geometric_median_objective(median: torch.Tensor, points: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (torch.linalg.norm(points - median.view(1, -1), dim=1) * weights).sum()


def compute_geometric_median(points: torch.Tensor, weights: Optional[torch.Tensor] = None, eps: float = 1e-6, maxiter: int = 100,
                             ftol: float = 1e-20, do_log: bool = False, **_ignored):
    """points [n, d] -> SimpleNamespace(median [d], new_weights, termination, logs)."""
    with torch.no_grad():
        if weights is None:
            weights = torch.ones((points.shape[0],), device=points.device)
        new_weights = weights
        median = weighted_average(points, weights)
        objective = geometric_median_objective(median, points, weights)
        logs = [objective] if do_log else None
        converged = False
        for _ in range(maxiter):
            previous = objective
            norms = torch.linalg.norm(points - median.view(1, -1), dim=1)
            new_weights = weights / torch.clamp(norms, min=eps)
            median = weighted_average(points, new_weights)
            objective = geometric_median_objective(median, points, weights)
            if logs is not None:
                logs.append(objective)
            if abs(previous - objective) <= ftol * objective:
                converged = True
                break
    median = weighted_average(points, new_weights)
    return SimpleNamespace(median=median, new_weights=new_weights, logs=logs,
                           termination="function value converged within tolerance" if converged else "maximum iterations reached")
