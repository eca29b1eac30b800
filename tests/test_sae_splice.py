"""HookedSAEViT (reference models/base_vit.py:827-1086): SAE splice bookkeeping on CPU, splice numerics on the GPU."""
import pytest
import torch

from vit_prisma.configs.HookedViTConfig import HookedViTConfig
from vit_prisma.models.base_vit import HookedSAEViT
from vit_prisma.prisma_tools.hook_point import HookPoint
from vit_prisma.sae.config import VisionModelSAERunnerConfig
from vit_prisma.sae.sae import StandardSparseAutoencoder

VIT = dict(n_layers=2, d_model=64, d_head=16, n_heads=4, d_mlp=128, patch_size=16, image_size=32, n_classes=10)


def _sae(device, layer=0, k=8):
    cfg = VisionModelSAERunnerConfig(d_in=64, expansion_factor=4, activation_fn_str="topk", activation_fn_kwargs={"k": k}, _device=device,
                                     _dtype="float32", hook_point_layer=layer, layer_subtype="hook_resid_post", log_to_wandb=False,
                                     n_checkpoints=0, checkpoint_path="/tmp/unused")
    torch.manual_seed(layer)
    return StandardSparseAutoencoder(cfg)


def test_splice_bookkeeping_on_cpu():
    model = HookedSAEViT(HookedViTConfig(**VIT))
    names_before = list(model.hook_dict)
    sae = _sae("cpu")
    assert sae.cfg.hook_point == "blocks.0.hook_resid_post"
    model.add_sae(sae)
    assert model.blocks[0].hook_resid_post is sae and model.acts_to_saes == {"blocks.0.hook_resid_post": sae}
    assert sae.cfg.return_out_only is True
    for inner in ("hook_sae_in", "hook_hidden_pre", "hook_hidden_post", "hook_sae_out"):
        assert f"blocks.0.hook_resid_post.{inner}" in model.hook_dict
    assert "blocks.0.hook_resid_post" not in model.hook_dict
    assert "SAE spliced in" in model._fused_blocker(torch.zeros(1))
    other = _sae("cpu")
    with model.saes(saes=[other], use_error_term=True):
        assert model.blocks[0].hook_resid_post is other and other.use_error_term is True
    assert model.blocks[0].hook_resid_post is sae and not hasattr(other, "_original_use_error_term") and other.use_error_term is False
    model.reset_saes()
    assert model.acts_to_saes == {} and isinstance(model.blocks[0].hook_resid_post, HookPoint)
    assert list(model.hook_dict) == names_before and model.hook_dict["blocks.0.hook_resid_post"].name == "blocks.0.hook_resid_post"
    with pytest.raises(ValueError):
        model.reset_saes(["a", "b"], [None])


@pytest.mark.gpu
def test_spliced_forward_equals_replacing_the_activation_by_hand():
    model = HookedSAEViT(HookedViTConfig(**VIT)).to("cuda").eval()
    sae = _sae("cuda", layer=1)
    sae.set_decoder_norm_to_unit_norm()
    x = torch.randn(3, 3, 32, 32, device="cuda")
    clean = model(x)
    assert model.last_route == "fused"
    spliced, cache = model.run_with_cache_with_saes(x, saes=[sae])
    assert model.last_route.startswith("hooked: SAE spliced in")
    name = "blocks.1.hook_resid_post"
    for inner in ("hook_sae_in", "hook_hidden_pre", "hook_hidden_post", "hook_sae_out"):
        assert f"{name}.{inner}" in cache
    assert name not in cache and cache[f"{name}.hook_hidden_post"].shape == (3, 5, 256)
    assert int((cache[f"{name}.hook_hidden_post"] > 0).sum(-1).max()) <= 8
    # the same computation spelled out with a replacing hook on the un-spliced model
    assert model.acts_to_saes == {} and isinstance(model.blocks[1].hook_resid_post, HookPoint)
    sae.cfg.return_out_only = True
    by_hand = model.run_with_hooks(x, fwd_hooks=[(name, lambda act, hook: sae(act))])
    assert torch.allclose(spliced, by_hand, rtol=1e-5, atol=1e-6)
    assert not torch.allclose(spliced, clean, rtol=1e-3, atol=1e-4), "an untrained SAE must change the output"
    # error term: the clean activation flows on, so the output equals the clean run while the SAE's hook points still fire
    out_err, cache_err = model.run_with_cache_with_saes(x, saes=[sae], use_error_term=True)
    assert torch.allclose(out_err, clean, rtol=1e-5, atol=1e-6) and f"{name}.hook_sae_out" in cache_err
    assert model(x) is not None and model.last_route == "fused"


@pytest.mark.gpu
def test_substitution_loss_matches_the_spliced_model_and_the_oracle_clean_loss():
    """get_substitution_loss (reference sae/evals/evals.py:321-391): clean loss vs the CPU oracle's forward, reconstruction loss vs the
    HookedSAEViT splice of the same SAE, zero-ablation loss vs a hand-rolled hook, score = (zero - recons) / (zero - clean)."""
    import torch.nn.functional as F
    from oracle.vit_oracle import vit_forward_with_cache
    from vit_prisma.sae.evals.evals import get_logits, get_similarity, get_substitution_loss, zero_ablate_hook
    torch.manual_seed(3)
    model = HookedSAEViT(HookedViTConfig(**VIT)).to("cuda").eval()
    sae = _sae("cuda", layer=1)
    sae.set_decoder_norm_to_unit_norm()
    x = torch.randn(6, 3, 32, 32)
    labels = torch.tensor([1, 4, 0, 9, 3, 3])
    out_dim = model(x.cuda()).shape[-1]                                    # the embedding the text features are compared with
    text = torch.nn.functional.normalize(torch.randn(10, out_dim), dim=-1)
    score, loss, recons, zero = get_substitution_loss(sae, model, x, labels, text, device=torch.device("cuda"))
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    cfg = dict(VIT, n_channels=3, eps=model.cfg.eps, activation_name=model.cfg.activation_name, normalization_type=model.cfg.normalization_type,
               use_cls_token=model.cfg.use_cls_token, layer_norm_pre=model.cfg.layer_norm_pre, normalize_output=model.cfg.normalize_output,
               return_type=model.cfg.return_type, classification_type=model.cfg.classification_type)
    ref_out, _ = vit_forward_with_cache(sd, cfg, x)
    ref_loss = F.cross_entropy(ref_out @ text.T, labels)
    assert abs(float(loss) - float(ref_loss)) <= 1e-4 * abs(float(ref_loss))
    spliced = model.run_with_saes(x.cuda(), saes=[sae])
    assert abs(float(recons) - float(F.cross_entropy(get_logits(spliced, text).float(), labels.cuda()))) <= 1e-5
    zeroed = model.run_with_hooks(x.cuda(), fwd_hooks=[("blocks.1.hook_resid_post", zero_ablate_hook)])
    assert abs(float(zero) - float(F.cross_entropy((zeroed.float() @ text.cuda().T), labels.cuda()))) <= 1e-5
    assert abs(float(score) - float((zero - recons) / (zero - loss))) <= 1e-6
    sm, top = get_similarity(spliced, text, k=3)
    assert sm.shape == (6, 10) and top.shape == (6, 3) and torch.allclose(sm.sum(-1), torch.ones(6, device="cuda"), atol=1e-5)
