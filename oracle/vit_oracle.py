"""ORACLE (test infrastructure, not product code): CPU restatement of HookedViT.run_with_cache.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module, and only as the *checker*.  The product path (vit_prisma -> libprisma_b200.so) never
touches it and has no CPU route of its own.

What is restated: the forward of the reference's HookedViT together with the ``_save_hook`` caching
semantics, as ONE plain function over a state dict -- no nn.Module, no HookPoint machinery -- in
plain PyTorch CPU arithmetic (fp32, or bf16 with the reference's rounding points).  Each step cites
the reference lines it follows (paths relative to /root/reference/src/vit_prisma).

Pinning: tests/test_oracle_golden.py checks this function against fixtures produced by running the
UNMODIFIED reference in the build container (tests/golden/make_golden.py):
  * vit_tiny_*.pt   -- every cache key, full tensors, two configs, fp32 and bf16;
  * vit_b32_fp32_digest.pt -- CLIP ViT-B/32 geometry, batch 4, per-key digests + full output.
The reference's own numeric tests for this path compare against open_clip/timm/HF checkpoints that
need network access, so the reference itself, run here, is the anchor.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F


def _act(name: str, x: torch.Tensor) -> torch.Tensor:
    # models/layers/mlp.py:41-52, models/activation_fns.py:19-47
    if name == "relu":
        return F.relu(x)
    if name == "gelu":
        return F.gelu(x)
    if name == "silu":
        return F.silu(x)
    if name == "gelu_new":
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))
    if name == "gelu_fast":
        return 0.5 * x * (1.0 + torch.tanh(x * 0.7978845608 * (1.0 + 0.044715 * x * x)))
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    raise ValueError(name)


def _layer_norm(x, w, b, eps, dtype, emit, prefix):
    # models/layers/layer_norm.py:75-93 (LayerNorm) / :27-45 (LayerNormPre when w is None)
    if dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float32)
    x = x - x.mean(-1, keepdim=True)
    scale = (x.pow(2).mean(-1, keepdim=True) + eps).sqrt()
    emit(prefix + "hook_scale", scale)
    y = x / scale
    if w is not None:
        y = y * w + b          # fp32 * model-dtype parameter -> fp32 (type promotion), as in the reference
    emit(prefix + "hook_normalized", y)
    return y.to(dtype)


def vit_forward_with_cache(sd: Dict[str, torch.Tensor], cfg: dict, images: torch.Tensor,
                           names_filter: Optional[Callable[[str], bool]] = None,
                           stop_at_layer: Optional[int] = None):
    """Returns (model_out, OrderedDict cache) exactly like
    ``HookedViT.run_with_cache(images, names_filter=..., stop_at_layer=..., return_cache_object=False)``.

    ``cfg`` holds the HookedViTConfig fields that shape the graph: n_layers, d_model, d_head, n_heads,
    d_mlp, patch_size, image_size, n_channels, n_classes, eps, activation_name, normalization_type,
    use_cls_token, layer_norm_pre, normalize_output, return_type, classification_type, dtype,
    use_attn_scale.
    """
    dtype = cfg.get("dtype", torch.float32)
    want = names_filter or (lambda name: True)
    cache: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def emit(name: str, t: torch.Tensor):
        # prisma_tools/hooked_root_module.py:312-316: cache[hook.name] = tensor.detach().to(device)
        if want(name):
            cache[name] = t.detach()
        return t

    L, d, H, dh = cfg["n_layers"], cfg["d_model"], cfg["n_heads"], cfg["d_head"]
    P = cfg["patch_size"]
    affine = cfg.get("normalization_type", "LN") == "LN"
    eps = cfg["eps"]
    B = images.shape[0]
    x = images.to(dtype)

    # models/layers/patch_embedding.py:26-32: Conv2d(k=s=P).flatten(2).transpose(1,2)
    embed = F.conv2d(x, sd["embed.proj.weight"], sd["embed.proj.bias"], stride=P).flatten(2).transpose(1, 2)
    emit("hook_embed", embed)
    if cfg.get("use_cls_token", True):
        # models/base_vit.py:171-175
        embed = torch.cat((sd["cls_token"].expand(B, -1, -1), embed), dim=1)
    # models/layers/position_embedding.py:32-38 (broadcast view of W_pos)
    pos = sd["pos_embed.W_pos"].unsqueeze(0).expand(B, -1, -1)
    emit("hook_pos_embed", pos)
    resid = embed + pos                                   # base_vit.py:179
    emit("hook_full_embed", resid)
    if cfg.get("layer_norm_pre", False):
        resid = _layer_norm(resid, sd.get("ln_pre.w") if affine else None, sd.get("ln_pre.b") if affine else None,
                            eps, dtype, emit, "ln_pre.")  # base_vit.py:183-185
        emit("hook_ln_pre", resid)

    attn_scale = math.sqrt(dh) if cfg.get("use_attn_scale", True) else 1.0
    blocks = range(L)[:stop_at_layer] if stop_at_layer is not None else range(L)
    for l in blocks:
        p = f"blocks.{l}."
        emit(p + "hook_resid_pre", resid)                 # layers/transformer_block.py:87
        # ln1 is applied to q, k and v inputs separately (:106-111); identical inputs -> identical outputs,
        # the cache keeps the last firing.
        n1 = _layer_norm(resid, sd.get(p + "ln1.w") if affine else None, sd.get(p + "ln1.b") if affine else None,
                         eps, dtype, emit, p + "ln1.")
        # layers/attention.py:186-244: einsum("b p d, h d e -> b p h e") + b
        q = torch.einsum("bpd,hde->bphe", n1, sd[p + "attn.W_Q"]) + sd[p + "attn.b_Q"]
        emit(p + "attn.hook_q", q)
        k = torch.einsum("bpd,hde->bphe", n1, sd[p + "attn.W_K"]) + sd[p + "attn.b_K"]
        emit(p + "attn.hook_k", k)
        v = torch.einsum("bpd,hde->bphe", n1, sd[p + "attn.W_V"]) + sd[p + "attn.b_V"]
        emit(p + "attn.hook_v", v)
        scores = torch.einsum("bqhe,bkhe->bhqk", q, k) / attn_scale   # attention.py:246-265
        emit(p + "attn.hook_attn_scores", scores)
        pattern = F.softmax(scores, dim=-1)                            # attention.py:148-150
        pattern = torch.where(torch.isnan(pattern), torch.zeros_like(pattern), pattern)
        emit(p + "attn.hook_pattern", pattern)
        pattern = pattern.to(dtype)
        z = torch.einsum("bkhe,bhqk->bqhe", v, pattern)                # attention.py:267-281
        emit(p + "attn.hook_z", z)
        attn_out = torch.einsum("bqhe,hed->bqd", z, sd[p + "attn.W_O"]) + sd[p + "attn.b_O"]   # :155-168
        emit(p + "hook_attn_out", attn_out)               # transformer_block.py:113-119 (dropout p=0)
        resid_mid = resid + attn_out                      # :121-124
        emit(p + "hook_resid_mid", resid_mid)
        n2 = _layer_norm(resid_mid, sd.get(p + "ln2.w") if affine else None, sd.get(p + "ln2.b") if affine else None,
                         eps, dtype, emit, p + "ln2.")
        pre = n2 @ sd[p + "mlp.W_in"] + sd[p + "mlp.b_in"]             # layers/mlp.py:67-70
        emit(p + "mlp.hook_pre", pre)
        post = _act(cfg["activation_name"], pre)                       # :71-72
        emit(p + "mlp.hook_post", post)
        mlp_out = post @ sd[p + "mlp.W_out"] + sd[p + "mlp.b_out"]     # :76-79
        emit(p + "hook_mlp_out", mlp_out)                 # transformer_block.py:130-133
        resid = resid_mid + mlp_out                       # :134
        emit(p + "hook_resid_post", resid)

    if stop_at_layer is not None:
        return resid, cache                               # base_vit.py:189-190

    xf = _layer_norm(resid, sd.get("ln_final.w") if affine else None, sd.get("ln_final.b") if affine else None,
                     eps, dtype, emit, "ln_final.")       # base_vit.py:192-193
    emit("hook_ln_final", xf)
    if cfg.get("classification_type", "cls") == "gaap":
        pooled = xf.mean(dim=1)                           # :195-198
    else:
        pooled = xf[:, 0]                                 # :199-208
    if cfg.get("return_type", "pre_logits") != "pre_logits":
        pooled = pooled @ sd["head.W_H"] + sd["head.b_H"]  # layers/head.py:27-38
    emit("hook_post_head_pre_normalize", pooled)          # base_vit.py:212
    if cfg.get("normalize_output", False):
        pooled = F.normalize(pooled, dim=-1)              # :214-215
    return pooled, cache


# ------------------------------------------------------------------------ helpers shared by tests / bench
CLIP_B32 = dict(n_layers=12, d_model=768, d_head=64, n_heads=12, d_mlp=3072, patch_size=32, image_size=224,
                n_channels=3, n_classes=512, eps=1e-5, activation_name="gelu", normalization_type="LN",
                use_cls_token=True, layer_norm_pre=True, normalize_output=True, return_type="class_logits",
                classification_type="cls")
CLIP_L14 = dict(n_layers=24, d_model=1024, d_head=64, n_heads=16, d_mlp=4096, patch_size=14, image_size=224,
                n_channels=3, n_classes=768, eps=1e-5, activation_name="gelu", normalization_type="LN",
                use_cls_token=True, layer_norm_pre=True, normalize_output=True, return_type="class_logits",
                classification_type="cls")


def state_dict_shapes(cfg: dict) -> Dict[str, tuple]:
    """Parameter names and shapes of a HookedViT with this config."""
    L, d, H, dh, M = cfg["n_layers"], cfg["d_model"], cfg["n_heads"], cfg["d_head"], cfg["d_mlp"]
    P, C = cfg["patch_size"], cfg.get("n_channels", 3)
    T = (cfg["image_size"] // P) ** 2 + (1 if cfg.get("use_cls_token", True) else 0)
    affine = cfg.get("normalization_type", "LN") == "LN"
    shapes: Dict[str, tuple] = {"cls_token": (1, 1, d), "embed.proj.weight": (d, C, P, P), "embed.proj.bias": (d,),
                                "pos_embed.W_pos": (T, d)}
    if cfg.get("layer_norm_pre", False) and affine:
        shapes["ln_pre.w"] = (d,)
        shapes["ln_pre.b"] = (d,)
    for l in range(L):
        p = f"blocks.{l}."
        if affine:
            for ln in ("ln1", "ln2"):
                shapes[p + ln + ".w"] = (d,)
                shapes[p + ln + ".b"] = (d,)
        for n in ("W_Q", "W_K", "W_V"):
            shapes[p + "attn." + n] = (H, d, dh)
        shapes[p + "attn.W_O"] = (H, dh, d)
        for n in ("b_Q", "b_K", "b_V"):
            shapes[p + "attn." + n] = (H, dh)
        shapes[p + "attn.b_O"] = (d,)
        shapes[p + "mlp.W_in"] = (d, M)
        shapes[p + "mlp.b_in"] = (M,)
        shapes[p + "mlp.W_out"] = (M, d)
        shapes[p + "mlp.b_out"] = (d,)
    if affine:
        shapes["ln_final.w"] = (d,)
        shapes["ln_final.b"] = (d,)
    shapes["head.W_H"] = (d, cfg["n_classes"])
    shapes["head.b_H"] = (cfg["n_classes"],)
    return shapes


def recipe_state_dict(shapes: Dict[str, tuple], seed: int = 1234, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (no checkpoint download is possible): the same recipe feeds the
    reference when fixtures are made and every implementation under test.  Scales keep activations O(1)
    through depth so parity errors are measured on realistic magnitudes."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in sorted(shapes):
        shape = shapes[name]
        r = torch.randn(shape, generator=g)
        leaf = name.split(".")[-1]
        if leaf == "w":                                   # LayerNorm gain
            t = 1.0 + 0.1 * r
        elif leaf in ("b", "bias", "b_Q", "b_K", "b_V", "b_O", "b_in", "b_out", "b_H"):
            t = 0.05 * r
        elif leaf == "cls_token" or name == "cls_token":
            t = 0.02 * r
        elif leaf == "W_pos":
            t = 0.02 * r
        elif leaf in ("W_Q", "W_K", "W_V"):
            t = r / math.sqrt(shape[1])
        elif leaf == "W_O":
            t = r / math.sqrt(shape[0] * shape[1])
        elif leaf == "weight":                            # conv patch embedding [d, C, P, P]
            t = r / math.sqrt(shape[1] * shape[2] * shape[3])
        else:                                             # W_in, W_out, W_H: [fan_in, fan_out]
            t = r / math.sqrt(shape[0])
        sd[name] = t.to(dtype)
    return sd


def digest(t: torch.Tensor, n_samples: int = 16) -> dict:
    """Small, order-sensitive fingerprint of a tensor for fixtures too large to commit."""
    tf = t.detach().to(torch.float64).reshape(-1)
    n = tf.numel()
    g = torch.Generator().manual_seed(n % 9973 + 17)
    idx = torch.randint(0, n, (min(n_samples, n),), generator=g)
    weights = torch.cos(torch.arange(n, dtype=torch.float64) * 0.37)
    return {"shape": tuple(t.shape), "dtype": str(t.dtype), "sum": float(tf.sum()), "abs_sum": float(tf.abs().sum()),
            "wsum": float((tf * weights).sum()), "idx": idx, "samples": tf[idx].to(torch.float32), "max_abs": float(tf.abs().max())}
