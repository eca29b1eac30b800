"""Learned position embedding (reference models/layers/position_embedding.py:13-38): returns ``W_pos``
broadcast over the batch as a stride-0 view -- no arithmetic, no copy."""
from __future__ import annotations

from typing import Dict, Union

import torch

from vit_prisma.b200.staging import host_staged
import torch.nn as nn

from vit_prisma.configs.HookedViTConfig import HookedViTConfig


class PosEmbedding(nn.Module):
    def __init__(self, cfg: Union[Dict, HookedViTConfig]):
        super().__init__()
        if isinstance(cfg, Dict):
            cfg = HookedViTConfig.from_dict(cfg)
        self.cfg = cfg
        n = (cfg.image_size // cfg.patch_size) ** 2
        if cfg.is_video_transformer:
            n *= cfg.video_num_frames // cfg.video_tubelet_depth
        self.W_pos = nn.Parameter(torch.empty(n + 1 if cfg.use_cls_token else n, cfg.d_model, dtype=cfg.dtype))

    @host_staged
    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        return self.W_pos.unsqueeze(0).expand(tokens.size(0), -1, -1)
