"""Seeded synthetic geometries, weights and activation pools for benchmarks and smoke runs.

There is no network on the build or GPU boxes, so neither pretrained CLIP checkpoints nor ImageNet are available:
``bench.py`` and ``__graft_entry__.smoke()`` run the product on the *architecture* BASELINE.json names with deterministic
synthetic weights and inputs.  The same recipe (same seed, same order of ``randn`` draws) is restated in ``oracle/`` for the
CPU checker; ``tests/test_host_logic.py`` holds the two copies to each other bit for bit.  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

# CLIP ViT-B/32 (DataComp-XL) and ViT-L/14 geometries (reference model_config_registry.py:43-57, 209-214)
CLIP_B32 = dict(n_layers=12, d_model=768, d_head=64, n_heads=12, d_mlp=3072, patch_size=32, image_size=224,
                n_channels=3, n_classes=512, eps=1e-5, activation_name="gelu", normalization_type="LN",
                use_cls_token=True, layer_norm_pre=True, normalize_output=True, return_type="class_logits",
                classification_type="cls")
CLIP_L14 = dict(CLIP_B32, n_layers=24, d_model=1024, n_heads=16, d_mlp=4096, patch_size=14, n_classes=768)

_BIAS_LEAVES = ("b", "bias", "b_Q", "b_K", "b_V", "b_O", "b_in", "b_out", "b_H")


def recipe_state_dict(shapes: Dict[str, tuple], seed: int = 1234, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic weights for a HookedViT state dict: one ``randn`` per parameter in sorted-name order, scaled by role so the
    residual stream stays O(1) through depth (parity errors are then measured on realistic magnitudes)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        r = torch.randn(shape, generator=g)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "w":
            t = 1.0 + 0.1 * r                                      # LayerNorm gain
        elif leaf in _BIAS_LEAVES:
            t = 0.05 * r
        elif leaf in ("cls_token", "W_pos"):
            t = 0.02 * r
        elif leaf in ("W_Q", "W_K", "W_V"):
            t = r / math.sqrt(shape[1])                            # [H, d_model, d_head]
        elif leaf == "W_O":
            t = r / math.sqrt(shape[0] * shape[1])                 # [H, d_head, d_model]
        elif leaf == "weight":
            t = r / math.sqrt(shape[1] * shape[2] * shape[3])      # conv patch embedding [d, C, P, P]
        else:
            t = r / math.sqrt(shape[0])                            # W_in, W_out, W_H: [fan_in, fan_out]
        out[name] = t.to(dtype)
    return out


def activation_pool(tokens: int, d: int, seed: int = 0) -> torch.Tensor:
    """Residual-stream-like activations: ``randn * 2 + per-column offset`` -- the non-zero mean makes the b_dec
    initialisation, the run-time layer norm and the batch centring of the loss all matter."""
    g = torch.Generator().manual_seed(seed)
    off = torch.randn(d, generator=g)
    return torch.randn(tokens, d, generator=g) * 2.0 + off


def sae_init_params(d: int, F: int, seed: int = 0, device="cpu") -> Dict[str, torch.Tensor]:
    """Unit-norm decoder rows and an encoder whose [d, F] rows are unit-norm (the reference's ``independent`` init geometry,
    sae.py:104-130, 537-555), from one seeded generator so every rank / arm starts from the same dictionary."""
    g = torch.Generator().manual_seed(seed)
    W_dec = torch.randn(F, d, generator=g)
    W_dec /= W_dec.norm(dim=1, keepdim=True)
    W_encT = torch.randn(F, d, generator=g)
    W_encT /= W_encT.norm(dim=0, keepdim=True) + 1e-12
    return dict(W_encT=W_encT.to(device), W_dec=W_dec.to(device), b_enc=torch.zeros(F, device=device), b_dec=torch.zeros(d, device=device))
