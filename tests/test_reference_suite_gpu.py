"""The reference's own offline tests, re-expressed against this package on the GPU
(reference tests/test_hooks.py:33-231, tests/test_cache_hook_names.py:22-55, tests/models/test_models.py:23-83).
Same assertions; the model and input live on cuda because the hot path has no CPU route."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    from vit_prisma.configs.HookedViTConfig import HookedViTConfig
    from vit_prisma.models.base_vit import HookedViT
    # positional order (n_layers, d_model, d_head, d_mlp) as in the reference test
    return HookedViT(HookedViTConfig(1, 8, 8, 8, return_type="logits")).to("cuda")


@pytest.fixture(scope="module")
def image():
    return torch.rand((2, 3, 224, 224), device="cuda")


embed = lambda name: name == "hook_embed"  # noqa: E731


class Counter:
    def __init__(self):
        self.count = 0

    def inc(self, *args, **kwargs):
        self.count += 1


def n_embed_hooks(model):
    return len(model.hook_dict["hook_embed"].fwd_hooks)


def test_hook_attaches_normally(model, image):
    c = Counter()
    model.run_with_hooks(image, fwd_hooks=[(embed, c.inc)])
    assert all(len(hp.fwd_hooks) == 0 for hp in model.hook_dict.values())
    assert c.count == 1
    model.remove_all_hook_fns(including_permanent=True)


def test_perma_hook_attaches_normally(model, image):
    c = Counter()
    model.add_perma_hook(embed, c.inc)
    assert n_embed_hooks(model) == 1
    model.run_with_hooks(image, fwd_hooks=[])
    assert n_embed_hooks(model) == 1 and c.count == 1
    model.remove_all_hook_fns(including_permanent=True)


def test_nested_hook_context_manager(model, image):
    c = Counter()
    with model.hooks(fwd_hooks=[(embed, c.inc)]):
        assert n_embed_hooks(model) == 1
        model.forward(image)
        assert c.count == 1
        with model.hooks(fwd_hooks=[(embed, c.inc)]):
            assert n_embed_hooks(model) == 2
            model.forward(image)
            assert c.count == 3
        assert n_embed_hooks(model) == 1
    assert n_embed_hooks(model) == 0 and c.count == 3
    model.remove_all_hook_fns(including_permanent=True)


def test_context_manager_run_with_cache(model, image):
    c = Counter()
    with model.hooks(fwd_hooks=[(embed, c.inc)]):
        assert n_embed_hooks(model) == 1
        model.run_with_cache(image)
        assert model.last_route.startswith("hooked")
        assert n_embed_hooks(model) == 1
    assert n_embed_hooks(model) == 0 and c.count == 1
    model.remove_all_hook_fns(including_permanent=True)


def test_hook_context_manager_with_permanent_hook(model, image):
    c = Counter()
    model.add_perma_hook(embed, c.inc)
    with model.hooks(fwd_hooks=[(embed, c.inc)]):
        assert n_embed_hooks(model) == 2
        model.forward(image)
    assert n_embed_hooks(model) == 1 and c.count == 2
    model.remove_all_hook_fns(including_permanent=True)


def test_nested_context_manager_with_failure(model, image):
    def fail_hook(z, hook):
        raise ValueError("fail")

    c = Counter()
    with model.hooks(fwd_hooks=[(embed, c.inc)]):
        with pytest.raises(ValueError):
            with model.hooks(fwd_hooks=[(embed, fail_hook)]):
                assert n_embed_hooks(model) == 2
                model.forward(image)
        assert n_embed_hooks(model) == 1 and c.count == 1
    assert n_embed_hooks(model) == 0
    model.remove_all_hook_fns(including_permanent=True)


def test_remove_hook(model, image):
    c = Counter()
    model.add_perma_hook(embed, c.inc)
    model.remove_all_hook_fns()
    assert n_embed_hooks(model) == 1
    model.remove_all_hook_fns(including_permanent=True)
    assert n_embed_hooks(model) == 0
    model.run_with_hooks(image, fwd_hooks=[])
    assert c.count == 0


def test_conditional_hooks(model, image):
    def identity_hook(z, hook):
        return z

    for hook_name, setter in [("blocks.0.attn.hook_result", model.set_use_attn_result),
                              ("blocks.0.hook_q_input", model.set_use_split_qkv_input),
                              ("blocks.0.hook_mlp_in", model.set_use_hook_mlp_in),
                              ("blocks.0.hook_attn_in", model.set_use_attn_in)]:
        model.reset_hooks()
        setter(False)
        with pytest.raises(AssertionError):
            model.add_hook(hook_name, identity_hook)
        setter(True)
        model.add_hook(hook_name, identity_hook)
        setter(False)
    shapes = {3: (2, 50, model.cfg.d_model), 4: (2, 50, model.cfg.n_heads, model.cfg.d_model)}
    for hook_name, setter, nd in [("blocks.0.hook_q_input", model.set_use_split_qkv_input, 4),
                                  ("blocks.0.hook_attn_in", model.set_use_attn_in, 4),
                                  ("blocks.0.hook_mlp_in", model.set_use_hook_mlp_in, 3)]:
        model.reset_hooks()
        setter(True)
        cache = model.run_with_cache(image, names_filter=lambda x: x == hook_name)[1]
        assert list(cache.keys()) == [hook_name]
        assert cache[hook_name].shape == shapes[nd]
        setter(False)
    model.reset_hooks()


def test_attn_result_route_matches_default(model, image):
    """use_attn_result exposes per-head results; their head-sum must reproduce the default output."""
    model.reset_hooks()
    base = model(image)
    model.set_use_attn_result(True)
    out, cache = model.run_with_cache(image, names_filter="blocks.0.attn.hook_result")
    model.set_use_attn_result(False)
    assert cache["blocks.0.attn.hook_result"].shape == (2, 50, model.cfg.n_heads, model.cfg.d_model)
    assert torch.allclose(out, base, atol=1e-5)


@pytest.mark.parametrize("zero_attach_pos,prepend", [(z, p) for z in range(2) for p in [True, False]])
def test_prepending_hooks(model, image, zero_attach_pos, prepend):
    def set_to_zero(z, hook):
        z[:] = 0.0
        return z

    def set_to_randn(z, hook):
        return torch.randn_like(z) * 0.1

    model.reset_hooks()
    for hook_idx in range(2):
        model.add_hook("blocks.0.hook_resid_post", set_to_zero if hook_idx == zero_attach_pos else set_to_randn, prepend=prepend)
    logits = model(image[0][None, ...])
    logits_are_unembed_bias = (zero_attach_pos == 1) != prepend
    assert torch.allclose(logits, model.head.b_H[None, :]) == logits_are_unembed_bias
    model.reset_hooks()


def test_cache_hook_names_solu_ln(image):
    from vit_prisma.configs.HookedViTConfig import HookedViTConfig
    from vit_prisma.models.base_vit import HookedViT
    m = HookedViT(HookedViTConfig(1, 8, 8, 8, return_type="logits", activation_name="solu_ln")).to("cuda")
    expected = ["hook_embed", "hook_pos_embed", "hook_full_embed", "blocks.0.hook_resid_pre", "blocks.0.ln1.hook_scale",
                "blocks.0.ln1.hook_normalized", "blocks.0.attn.hook_q", "blocks.0.attn.hook_k", "blocks.0.attn.hook_v",
                "blocks.0.attn.hook_attn_scores", "blocks.0.attn.hook_pattern", "blocks.0.attn.hook_z", "blocks.0.hook_attn_out",
                "blocks.0.hook_resid_mid", "blocks.0.ln2.hook_scale", "blocks.0.ln2.hook_normalized", "blocks.0.mlp.hook_pre",
                "blocks.0.mlp.hook_mid", "blocks.0.mlp.ln.hook_scale", "blocks.0.mlp.ln.hook_normalized", "blocks.0.mlp.hook_post",
                "blocks.0.hook_mlp_out", "blocks.0.hook_resid_post", "ln_final.hook_scale", "ln_final.hook_normalized",
                "hook_ln_final", "hook_post_head_pre_normalize"]
    _, cache = m.run_with_cache(image)
    assert list(cache.keys()) == expected


def test_layer_shapes():
    from vit_prisma.configs.HookedViTConfig import HookedViTConfig
    from vit_prisma.models.base_vit import HookedViT
    from vit_prisma.models.layers.attention import Attention
    from vit_prisma.models.layers.mlp import MLP
    from vit_prisma.models.layers.patch_embedding import PatchEmbedding
    from vit_prisma.models.layers.transformer_block import TransformerBlock
    cfg = HookedViTConfig(n_layers=1, d_head=8, d_model=8, d_mlp=8)
    x = torch.randn(8, 16, cfg.d_model, device="cuda")
    HookedViT(cfg)  # init params of sub-layers the way the model does
    for mod in (Attention(cfg), MLP(cfg), TransformerBlock(cfg)):
        for p in mod.parameters():
            torch.nn.init.normal_(p, std=0.1)
        mod = mod.cuda()
        out = mod(x, x, x) if isinstance(mod, Attention) else mod(x)
        assert out.shape == (8, 16, cfg.d_model)
    cfg.return_type = "class_logits"
    m = HookedViT(cfg).cuda()
    img = torch.randn(8, cfg.n_channels, cfg.image_size, cfg.image_size, device="cuda")
    assert m(img).shape == (8, cfg.n_classes)
    pe = PatchEmbedding(cfg).cuda()
    assert pe(img).shape == (8, (cfg.image_size // cfg.patch_size) ** 2, cfg.d_model)
