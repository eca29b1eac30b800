"""Replacement-hook evaluation of a trained sparse coder inside the ViT (reference sae/evals/evals.py:321-404, 436-441): the joint
caller of both hot paths -- ``HookedViT.run_with_cache`` / ``run_with_hooks`` with the SAE forward spliced in at its hook point.

Only the compute pieces of the reference module live here (dataset loading, plotting and the text-label download need network /
plotting packages and are out of scope).  The ViT and SAE forwards run on the package's CUDA kernels; the class
logits are one small GEMM (``image_features @ text_features.T``); softmax / cross-entropy over [batch, n_classes] are scalar-sized
glue on the device."""
from __future__ import annotations

from functools import partial
from typing import Any

import torch
import torch.nn.functional as F

from vit_prisma.b200 import ops


def get_logits(image_features: torch.Tensor, text_features: torch.Tensor, device="cuda") -> torch.Tensor:
    """``image_features @ text_features.T`` (reference :394-395); text_features [n_classes, d] is already the K-major operand."""
    img = image_features.to(device).contiguous()
    txt = text_features.to(device, img.dtype).contiguous()
    out, _ = ops.gemm(img, txt, None)
    return out


def get_similarity(image_features: torch.Tensor, text_features: torch.Tensor, k: int = 5, device="cuda"):
    softmax_values = get_logits(image_features, text_features, device).float().softmax(dim=-1)     # reference :398-404
    _top_values, top_k_indices = torch.topk(softmax_values, k, dim=-1)
    return softmax_values, top_k_indices


def zero_ablate_hook(activations: torch.Tensor, hook: Any):
    return torch.zeros_like(activations)                                                        # reference :436-438


def get_feature_probability(feature_acts: torch.Tensor) -> torch.Tensor:
    return (feature_acts.abs() > 0).float().flatten(0, 1)                                       # reference :440-441


@torch.no_grad()
def get_substitution_loss(sparse_autoencoder, model, batch_tokens: torch.Tensor, gt_labels: torch.Tensor, text_embeddings: torch.Tensor,
                          device: torch.device = torch.device("cuda")):
    """(score, loss, recons_loss, zero_abl_loss): zero-shot cross-entropy of the clean model, of the model with the hook point's
    activation replaced by the SAE reconstruction, and with it zero-ablated; score = (zero_abl - recons) / (zero_abl - clean)
    (reference :321-391, including its per-head variant when ``cfg.hook_point_head_index`` is set)."""
    model = model.to(device)
    batch_tokens, gt_labels, text_embeddings = batch_tokens.to(device), gt_labels.to(device), text_embeddings.to(device)
    image_embeddings, _ = model.run_with_cache(batch_tokens)
    loss = F.cross_entropy(get_logits(image_embeddings, text_embeddings, device=device).float(), gt_labels)
    head_index = sparse_autoencoder.cfg.hook_point_head_index
    hook_point = sparse_autoencoder.cfg.hook_point

    def standard_replacement_hook(activations: torch.Tensor, hook: Any):
        return sparse_autoencoder.forward(activations)[0].to(activations.dtype)

    def head_replacement_hook(activations: torch.Tensor, hook: Any):
        new_activations = sparse_autoencoder.forward(activations[:, :, head_index].contiguous())[0].to(activations.dtype)
        activations[:, :, head_index] = new_activations
        return activations

    replacement_hook = standard_replacement_hook if head_index is None else head_replacement_hook
    recons_image_embeddings = model.run_with_hooks(batch_tokens, fwd_hooks=[(hook_point, partial(replacement_hook))])
    recons_loss = F.cross_entropy(get_logits(recons_image_embeddings, text_embeddings, device=device).float(), gt_labels)
    zero_abl_image_embeddings = model.run_with_hooks(batch_tokens, fwd_hooks=[(hook_point, zero_ablate_hook)])
    zero_abl_loss = F.cross_entropy(get_logits(zero_abl_image_embeddings, text_embeddings, device=device).float(), gt_labels)
    score = (zero_abl_loss - recons_loss) / (zero_abl_loss - loss)
    return score, loss, recons_loss, zero_abl_loss
