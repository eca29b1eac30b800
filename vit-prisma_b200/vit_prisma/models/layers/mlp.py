"""Hooked MLP (reference models/layers/mlp.py:15-80): ``hook_pre`` -> act -> ``hook_post`` -> out.

Hooked path: GEMM(+bias) -> hook_pre -> activation kernel -> hook_post -> GEMM(+bias).  When no hook is
attached to ``hook_pre``/``hook_post`` the activation rides in the first GEMM's epilogue (dual output).
"""
from __future__ import annotations

from typing import Dict, Union

import torch

from vit_prisma.b200.staging import host_staged
import torch.nn as nn

from vit_prisma.b200 import ops
from vit_prisma.b200.packing import PackCache, pack_t, with_lo
from vit_prisma.configs.HookedViTConfig import HookedViTConfig
from vit_prisma.models import activation_fns
from vit_prisma.models.layers.layer_norm import LayerNorm, LayerNormPre
from vit_prisma.prisma_tools.hook_point import HookPoint


class MLP(nn.Module):
    def __init__(self, cfg: Union[Dict, HookedViTConfig]):
        super().__init__()
        if isinstance(cfg, Dict):
            cfg = HookedViTConfig.from_dict(cfg)
        self.cfg = cfg
        dt = cfg.dtype
        self.W_in = nn.Parameter(torch.empty(cfg.d_model, cfg.d_mlp, dtype=dt))
        self.b_in = nn.Parameter(torch.empty(cfg.d_mlp, dtype=dt))
        self.W_out = nn.Parameter(torch.empty(cfg.d_mlp, cfg.d_model, dtype=dt))
        self.b_out = nn.Parameter(torch.empty(cfg.d_model, dtype=dt))
        self.hook_pre = HookPoint()    # [batch, pos, d_mlp]
        self.hook_post = HookPoint()   # [batch, pos, d_mlp]

        name = cfg.activation_name
        if name not in activation_fns.BY_NAME:
            raise ValueError(f"Invalid activation function name: {name}")
        self.act_fn = activation_fns.BY_NAME[name]
        if name == "solu_ln":
            self.hook_mid = HookPoint()  # between solu and its LayerNorm
            self.ln = LayerNorm(cfg, cfg.d_mlp) if cfg.normalization_type == "LN" else LayerNormPre(cfg)
        self._packs = PackCache()

    def packed_in(self):
        return self._packs.get("win", (self.W_in,), lambda: with_lo(pack_t(self.W_in)))

    def packed_out(self):
        return self._packs.get("wout", (self.W_out,), lambda: with_lo(pack_t(self.W_out)))

    @host_staged
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        win, _ = self.packed_in()
        wout, _ = self.packed_out()
        name = self.cfg.activation_name
        fusable = name in activation_fns.ELEMENTWISE and self.hook_pre.is_inert and self.hook_post.is_inert
        if fusable:
            _, post = ops.gemm(x, win, self.b_in, act=name, want_pre=False, want_post=True)
        else:
            pre, _ = ops.gemm(x, win, self.b_in)
            pre = self.hook_pre(pre)
            if not name.endswith("_ln"):
                post = self.hook_post(self.act_fn(pre))
            else:
                mid = self.hook_mid(self.act_fn(pre))
                post = self.hook_post(self.ln(mid))
        out, _ = ops.gemm(post, wout, self.b_out)
        return out
