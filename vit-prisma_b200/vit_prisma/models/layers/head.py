"""Classification / projection head (reference models/layers/head.py:12-38): ``x @ W_H + b_H``."""
from __future__ import annotations

from typing import Dict, Union

import torch

from vit_prisma.b200.staging import host_staged
import torch.nn as nn

from vit_prisma.b200 import ops
from vit_prisma.b200.packing import PackCache, pack_t, with_lo
from vit_prisma.configs.HookedViTConfig import HookedViTConfig


class Head(nn.Module):
    def __init__(self, cfg: Union[Dict, HookedViTConfig]):
        super().__init__()
        if isinstance(cfg, Dict):
            cfg = HookedViTConfig.from_dict(cfg)
        self.cfg = cfg
        self.W_H = nn.Parameter(torch.empty(cfg.d_model, cfg.n_classes, dtype=cfg.dtype))
        self.b_H = nn.Parameter(torch.zeros(cfg.n_classes, dtype=cfg.dtype))
        self._packs = PackCache()

    def packed(self):
        return self._packs.get("wh", (self.W_H,), lambda: with_lo(pack_t(self.W_H)))

    @host_staged
    def forward(self, residual: torch.Tensor) -> torch.Tensor:
        w, _ = self.packed()
        out, _ = ops.gemm(residual, w, self.b_H)
        return out
