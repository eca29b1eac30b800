"""Small host-side helpers that the cache container depends on.

``get_act_name`` follows the shorthand grammar of the reference
(src/vit_prisma/utils/prisma_utils.py:202-302): ``get_act_name('k', 6)`` ->
``blocks.6.attn.hook_k``, ``'scale4ln1'`` -> ``blocks.4.ln1.hook_scale``, full
names pass through.  ``transpose`` swaps the last two dims (used by the weight
post-processing helpers).  ``Slice`` is the small indexing helper the analysis
methods of ActivationCache accept.
"""
from __future__ import annotations

import re
from typing import Optional, Tuple, Union

import numpy as np
import torch

_LAYER_TYPE_ALIAS = {"a": "attn", "m": "mlp", "b": "", "block": "", "blocks": "", "attention": "attn"}
_ACT_ALIAS = {
    "attn": "pattern", "attn_logits": "attn_scores", "key": "k", "query": "q", "value": "v",
    "mlp_pre": "pre", "mlp_mid": "mid", "mlp_post": "post",
}
_ATTN_ACTS = {"k", "v", "q", "z", "rot_k", "rot_q", "result", "pattern", "attn_scores"}
_MLP_ACTS = {"pre", "post", "mid", "pre_linear"}
_SHORTHAND = re.compile(r"([a-z]+)(\d+)([a-z]?.*)")


def get_act_name(name: str, layer: Optional[Union[int, str]] = None,
                 layer_type: Optional[str] = None) -> str:
    if ("." in name or name.startswith("hook_")) and layer is None and layer_type is None:
        return name  # already a full hook name
    packed = _SHORTHAND.match(name)
    if packed is not None:
        name, layer, layer_type = packed.groups(0)
    name = _ACT_ALIAS.get(name, name)

    if name in _ATTN_ACTS:
        layer_type = "attn"
    elif name in _MLP_ACTS:
        layer_type = "mlp"
    elif layer_type in _LAYER_TYPE_ALIAS:
        layer_type = _LAYER_TYPE_ALIAS[layer_type]

    parts = []
    if layer is not None:
        parts.append(f"blocks.{layer}")
    if layer_type:
        parts.append(str(layer_type))
    parts.append(f"hook_{name}")
    full = ".".join(parts)
    if name in ("scale", "normalized") and layer is None:
        full = f"ln_final.{full}"
    return full


def transpose(tensor: torch.Tensor) -> torch.Tensor:
    return tensor.transpose(-1, -2)


def to_numpy(x):
    if isinstance(x, np.ndarray):
        return x
    if isinstance(x, (list, tuple)):
        return np.array(x)
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, (int, float, bool, str)):
        return np.array(x)
    raise ValueError(f"Input to to_numpy has invalid type: {type(x)}")


SliceInput = Optional[Union[int, Tuple[int, ...], list, torch.Tensor, np.ndarray]]


class Slice:
    """Index helper: None -> everything, int -> one index (dim dropped),
    tuple -> python slice, list/array/tensor -> gather."""

    def __init__(self, input_slice: SliceInput = None):
        if isinstance(input_slice, tuple):
            self.slice, self.mode = slice(*input_slice), "slice"
        elif isinstance(input_slice, int):
            self.slice, self.mode = input_slice, "int"
        elif isinstance(input_slice, slice):
            self.slice, self.mode = input_slice, "slice"
        elif isinstance(input_slice, (list, torch.Tensor, np.ndarray)):
            self.slice, self.mode = to_numpy(input_slice), "array"
        elif input_slice is None:
            self.slice, self.mode = slice(None), "identity"
        else:
            raise ValueError(f"Invalid input_slice {input_slice}")

    def apply(self, tensor: torch.Tensor, dim: int = 0) -> torch.Tensor:
        index = [slice(None)] * tensor.ndim
        index[dim] = self.slice
        return tensor[tuple(index)]

    def indices(self, max_ctx: Optional[int] = None):
        if self.mode == "int":
            return np.array([self.slice], dtype=np.int64)
        if max_ctx is None:
            raise ValueError("max_ctx must be specified if slice is not an integer")
        return np.arange(max_ctx, dtype=np.int64)[self.slice]

    def __repr__(self) -> str:
        return f"Slice: {self.slice} Mode: {self.mode} "
