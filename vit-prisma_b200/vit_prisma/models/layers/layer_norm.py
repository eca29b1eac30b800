"""LayerNorm / LayerNormPre with ``hook_scale`` and ``hook_normalized``
(reference models/layers/layer_norm.py:10-93).

No hook on ``hook_scale``: one fused kernel (csrc/layernorm.cu) produces scale, the hooked fp32 result
and the model-dtype output.  With hooks on ``hook_scale`` the op is split in two launches -- scale,
*hook*, normalise-with-the-hooked-scale -- so a hook that replaces or edits the scale in place feeds
the division exactly as in the reference graph.
"""
from __future__ import annotations

from typing import Dict, Optional, Union

import torch

from vit_prisma.b200.staging import host_staged
import torch.nn as nn

from vit_prisma.b200 import ops
from vit_prisma.configs.HookedViTConfig import HookedViTConfig
from vit_prisma.prisma_tools.hook_point import HookPoint


class _HookedNorm(nn.Module):
    has_affine = False

    def _setup(self, cfg):
        if isinstance(cfg, Dict):
            cfg = HookedViTConfig.from_dict(cfg)
        self.cfg = cfg
        self.eps = cfg.eps
        self.hook_scale = HookPoint()        # [batch, pos, 1]
        self.hook_normalized = HookPoint()   # [batch, pos, length]

    @host_staged
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        w = self.w if self.has_affine else None
        b = self.b if self.has_affine else None
        if self.hook_scale.is_inert:
            _, normalized, out = ops.layernorm(x, w, b, self.eps, self.cfg.dtype, want_scale=False)
        else:
            scale = self.hook_scale(ops.layernorm_scale(x, self.eps))
            _, normalized, out = ops.layernorm(x, w, b, self.eps, self.cfg.dtype, want_scale=False, scale_in=scale)
        hooked = self.hook_normalized(normalized)
        if hooked is normalized and self.hook_normalized.is_inert:
            return out
        # a hook saw (and may have edited / replaced) the fp32-or-model-dtype tensor: cast what it left us
        return ops.cast(hooked.contiguous(), self.cfg.dtype)


class LayerNormPre(_HookedNorm):
    """Centre + normalise only (weights folded elsewhere)."""

    def __init__(self, cfg: Union[Dict, HookedViTConfig]):
        super().__init__()
        self._setup(cfg)


class LayerNorm(_HookedNorm):
    has_affine = True

    def __init__(self, cfg: Union[Dict, HookedViTConfig], length: Optional[int] = None):
        super().__init__()
        self._setup(cfg)
        self.length = self.cfg.d_model if length is None else length
        self.w = nn.Parameter(torch.ones(self.length, dtype=self.cfg.dtype))
        self.b = nn.Parameter(torch.zeros(self.length, dtype=self.cfg.dtype))
