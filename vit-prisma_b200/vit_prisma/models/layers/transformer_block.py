"""Pre-LN transformer block with the reference's ten block-level hook points
(reference models/layers/transformer_block.py:30-138).

Quirks kept on purpose because caches and tests observe them:
  * ``ln1`` runs once per q/k/v input (three firings of ln1.hook_scale / hook_normalized per block;
    the cache keeps the last one) -- :106-111;
  * ``hook_attn_in`` / ``hook_q_input`` / ``hook_k_input`` / ``hook_v_input`` / ``hook_mlp_in`` only
    fire when the matching ``cfg.use_*`` toggle is on, and then see a *copy* of the residual with a
    head axis ([B,T,H,d]) -- :93-104, :125-129.
When ln1 has no hooks attached the three applications are numerically identical, so it is launched
once and the result shared.
"""
from __future__ import annotations

from typing import Dict, Optional, Union

import torch

from vit_prisma.b200.staging import host_staged
import torch.nn as nn

from vit_prisma.b200 import ops
from vit_prisma.configs.HookedViTConfig import HookedViTConfig
from vit_prisma.models.layers.attention import Attention
from vit_prisma.models.layers.layer_norm import LayerNorm, LayerNormPre
from vit_prisma.models.layers.mlp import MLP
from vit_prisma.prisma_tools.hook_point import HookPoint


def add_head_dimension(tensor: torch.Tensor, n_heads: int, clone_tensor: bool = True) -> torch.Tensor:
    """[B,T,d] -> [B,T,H,d]; a stride-0 view unless ``clone_tensor``."""
    expanded = tensor.unsqueeze(2).expand(-1, -1, n_heads, -1)
    return expanded.clone() if clone_tensor else expanded


def _make_norm(cfg):
    if cfg.normalization_type == "LN":
        return LayerNorm(cfg)
    if cfg.normalization_type == "LNPre":
        return LayerNormPre(cfg)
    if cfg.normalization_type is None:
        return nn.Identity()
    raise ValueError(f"Invalid normalization type: {cfg.normalization_type}")


class TransformerBlock(nn.Module):
    def __init__(self, cfg: Union[Dict, HookedViTConfig], block_index=None):
        super().__init__()
        if isinstance(cfg, Dict):
            cfg = HookedViTConfig.from_dict(cfg)
        self.cfg = cfg
        self.ln1 = _make_norm(cfg)
        if not cfg.attn_only:
            self.ln2 = _make_norm(cfg)
        self.attn = Attention(cfg)
        if not cfg.attn_only:
            self.mlp = MLP(cfg)

        self.hook_attn_in = HookPoint()
        self.hook_q_input = HookPoint()
        self.hook_k_input = HookPoint()
        self.hook_v_input = HookPoint()
        self.hook_mlp_in = HookPoint()
        self.hook_attn_out = HookPoint()
        self.hook_mlp_out = HookPoint()
        self.hook_resid_pre = HookPoint()
        if not cfg.attn_only:
            self.hook_resid_mid = HookPoint()
        self.hook_resid_post = HookPoint()

        self.attn_dropout = nn.Dropout(cfg.attn_dropout_rate)
        self.mlp_dropout = nn.Dropout(cfg.mlp_dropout_rate)

    def _ln1_is_silent(self) -> bool:
        ln = self.ln1
        if isinstance(ln, nn.Identity):
            return True
        return ln.hook_scale.is_inert and ln.hook_normalized.is_inert

    @host_staged
    def forward(self, resid_pre: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        cfg = self.cfg
        resid_pre = self.hook_resid_pre(resid_pre)

        attn_in = resid_pre
        if cfg.use_attn_in or cfg.use_split_qkv_input:
            attn_in = add_head_dimension(resid_pre, cfg.n_heads, clone_tensor=False)
        if cfg.use_attn_in:
            attn_in = self.hook_attn_in(attn_in.clone())
        if cfg.use_split_qkv_input:
            q_in = self.hook_q_input(attn_in.clone())
            k_in = self.hook_k_input(attn_in.clone())
            v_in = self.hook_v_input(attn_in.clone())
        else:
            q_in = k_in = v_in = attn_in

        if q_in is k_in and k_in is v_in and self._ln1_is_silent():
            nq = nk = nv = self.ln1(q_in)
        else:
            nq, nk, nv = self.ln1(q_in), self.ln1(k_in), self.ln1(v_in)
        attn_out = self.attn(query_input=nq, key_input=nk, value_input=nv, attention_mask=attn_mask)
        attn_out = self.hook_attn_out(self.attn_dropout(attn_out))

        if cfg.attn_only:
            return self.hook_resid_post(ops.add(resid_pre, attn_out))

        resid_mid = self.hook_resid_mid(ops.add(resid_pre, attn_out))
        mlp_in = resid_mid if not cfg.use_hook_mlp_in else self.hook_mlp_in(resid_mid.clone())
        mlp_out = self.mlp(self.ln2(mlp_in))
        mlp_out = self.hook_mlp_out(self.mlp_dropout(mlp_out))
        return self.hook_resid_post(ops.add(resid_mid, mlp_out))


class BertBlock(TransformerBlock):
    """Post-LN variant (reference :141-246); same hook points, LayerNorm applied after each residual add."""

    @host_staged
    def forward(self, resid_pre: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        resid_pre = self.hook_resid_pre(resid_pre)
        attn_out = self.attn(query_input=resid_pre, key_input=resid_pre, value_input=resid_pre, attention_mask=attn_mask)
        attn_out = self.hook_attn_out(self.attn_dropout(attn_out))
        resid_mid = self.hook_resid_mid(ops.add(resid_pre, attn_out))
        normalized_mid = self.ln1(resid_mid)
        mlp_in = normalized_mid if not self.cfg.use_hook_mlp_in else self.hook_mlp_in(normalized_mid.clone())
        mlp_out = self.hook_mlp_out(self.mlp_dropout(self.mlp(mlp_in)))
        resid_post = self.hook_resid_post(ops.add(normalized_mid, mlp_out))
        return self.ln2(resid_post)
