"""CLIP ViT-L/14 geometry (cfg #4): T = 257 tokens, 24 layers, d_model 1024, 16 heads -- the attention kernel's
34-key-tile instantiation, 1024/4096-wide GEMMs -- against the pinned oracle on the same seeded weights and images."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.vit_oracle import CLIP_L14, recipe_state_dict, vit_forward_with_cache  # noqa: E402
from tests.util import assert_close, rel_err  # noqa: E402


def _model(dtype):
    from vit_prisma.configs.HookedViTConfig import HookedViTConfig
    from vit_prisma.models.base_vit import HookedViT
    model = HookedViT(HookedViTConfig(**CLIP_L14, dtype=dtype))
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = recipe_state_dict(shapes, 4321)
    model.load_state_dict(sd)
    return model.to("cuda", dtype).eval(), sd


def _images(batch, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, 224, 224, generator=g)


def test_l14_activation_store_call_matches_oracle_fp32():
    """names_filter=[blocks.l.hook_resid_post] + stop_at_layer=l+1, exactly what VisionActivationsStore.get_activations issues
    (activations_store.py:262-270)."""
    model, sd = _model(torch.float32)
    x = _images(3)
    layer = 9
    name = f"blocks.{layer}.hook_resid_post"
    with torch.no_grad():
        ref_out, ref_cache = vit_forward_with_cache(sd, dict(CLIP_L14), x, names_filter=lambda n: n == name, stop_at_layer=layer + 1)
    out, cache = model.run_with_cache(x.cuda(), names_filter=[name], stop_at_layer=layer + 1)
    assert list(cache.keys()) == [name] == list(ref_cache.keys())
    assert tuple(out.shape) == (3, 257, 1024)
    assert_close(cache[name].cpu(), ref_cache[name], 1e-4, name)
    assert_close(out.cpu(), ref_out, 1e-4, "stop_at_layer output")


def test_l14_all_hook_points_match_oracle_fp32():
    """Full depth, every default-firing hook point (418 keys for L/14), batch 1."""
    model, sd = _model(torch.float32)
    x = _images(1, seed=3)
    with torch.no_grad():
        ref_out, ref_cache = vit_forward_with_cache(sd, dict(CLIP_L14), x)
    out, cache = model.run_with_cache(x.cuda())
    assert model.last_route == "fused"
    assert list(cache.keys()) == list(ref_cache.keys()) and len(cache) == 418
    worst = max(((rel_err(cache[k].float().cpu(), ref_cache[k].float()), k) for k in ref_cache), key=lambda t: t[0])
    print(f"[l14 fp32] worst key {worst[1]} rel err {worst[0]:.2e}")
    assert worst[0] < 1e-4, worst
    assert_close(out.cpu(), ref_out, 1e-4, "model output")


def test_l14_bf16_resid_post_close_to_fp32_truth():
    """bf16 mode on the L/14 geometry: the cached residual stream of layer 9 within 2e-2 of the fp32 oracle (two bf16 ulps after
    ten rounded residual adds, the same bar as tests/test_vit_gpu.py)."""
    model, sd = _model(torch.bfloat16)
    x = _images(2, seed=5)
    name = "blocks.9.hook_resid_post"
    sd16 = {k: v.to(torch.bfloat16).float() for k, v in sd.items()}
    with torch.no_grad():
        _, ref_cache = vit_forward_with_cache(sd16, dict(CLIP_L14), x.to(torch.bfloat16).float(), names_filter=lambda n: n == name,
                                              stop_at_layer=10)
    _, cache = model.run_with_cache(x.cuda().to(torch.bfloat16), names_filter=[name], stop_at_layer=10)
    assert rel_err(cache[name].float().cpu(), ref_cache[name]) < 2e-2
