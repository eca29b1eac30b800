"""Parity at BASELINE.json's FULL sizes and across the store -> trainer seam ("parity hardening"):

* cfg #3: one training step at 4096 tokens x d_sae 24576 against the oracle (tile-quantisation, the m-fastest raster, the hot-feature
  queues and the fused encoder all behave differently here than on the 64-token fixtures);
* cfg #2: run_with_cache at batch 512, eight sampled hook points of four of the 512 images against the oracle;
* VisionActivationsStore driven by a real HookedViT: get_activations vs the oracle's cache, half-buffer mix bookkeeping.
"""
import contextlib
import io

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.sae_oracle import new_adam_state, sae_train_step  # noqa: E402
from oracle.vit_oracle import CLIP_B32, recipe_state_dict, state_dict_shapes, vit_forward_with_cache  # noqa: E402
from tests.util import assert_close, rel_err  # noqa: E402


def test_cfg3_full_size_train_step_matches_oracle():
    from vit_prisma.b200.sae_engine import SaeStepEngine, unit_norm_rows_
    from vit_prisma.b200.synthetic import activation_pool, sae_init_params
    d, F, k, rows, lr = 768, 24576, 32, 4096, 1e-3
    p0 = sae_init_params(d, F, seed=0)
    x = activation_pool(rows, d, seed=5)
    p0["b_dec"] = activation_pool(8192, d, seed=0).mean(0)
    p0["b_enc"] = 0.01 * torch.randn(F, generator=torch.Generator().manual_seed(1))
    p = {"W_enc": p0["W_encT"].t().contiguous(), "W_dec": p0["W_dec"].clone(), "b_enc": p0["b_enc"].clone(), "b_dec": p0["b_dec"].clone()}
    state = new_adam_state(p)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    ref = sae_train_step(p, state, x, k, lr, 1)                                    # in place on p: parameters after the step
    eng = SaeStepEngine(p0["W_encT"].cuda(), p0["W_dec"].cuda(), p0["b_enc"].cuda(), p0["b_dec"].cuda(), k=k)
    assert eng.encoder == "fused"
    unit_norm_rows_(eng.W_dec)
    eng.refresh_lo()
    since, freq = torch.zeros(F, device="cuda"), torch.zeros(F, device="cuda")
    eng.train_step(x.cuda(), lr, since_fired=since, act_freq=freq, want_out=True)
    sc = eng.scalars_dict()
    assert abs(sc["mse"] - float(ref["mse"])) <= 1e-4 * float(ref["mse"]), (sc["mse"], float(ref["mse"]))
    assert abs(sc["grad_norm"] - float(ref["grad_norm"])) <= 1e-4 * float(ref["grad_norm"]), (sc["grad_norm"], float(ref["grad_norm"]))
    same = (eng.idx.cpu().long() == ref["idx"]).all(dim=1)
    hp = ref["fwd"]["hidden_pre"]
    gap = torch.topk(hp, k + 1, dim=-1).values
    near = ((gap[:, :-1] - gap[:, 1:]).abs().min(dim=1).values < 2e-6 * hp.abs().max())
    assert bool((same | near).all()), f"{(~(same | near)).sum().item()} rows differ beyond fp32 near-ties"
    assert same.float().mean().item() >= 0.999
    assert eng.fallback_rows() <= 8, eng.fallback_rows()
    assert_close(eng.sae_out.cpu(), ref["fwd"]["sae_out"], 1e-4, "sae_out")
    ref_dec = p["W_dec"] / p["W_dec"].norm(dim=1, keepdim=True)
    assert_close(eng.W_dec.cpu(), ref_dec, 1e-4, "W_dec after the step")
    assert_close(eng.W_encT.t().cpu(), p["W_enc"], 1e-4, "W_enc after the step")
    assert_close(eng.b_enc.cpu(), p["b_enc"], 1e-4, "b_enc after the step")
    assert_close(eng.b_dec.cpu(), p["b_dec"], 1e-4, "b_dec after the step")
    fired = (ref["fwd"]["feature_acts"] > 0).float().sum(0)
    assert torch.equal(freq.cpu(), fired) or (freq.cpu() - fired).abs().sum() <= 2 * (~same).sum()


def test_cfg2_batch_512_sampled_keys_match_oracle():
    from vit_prisma.configs.HookedViTConfig import HookedViTConfig
    from vit_prisma.models.base_vit import HookedViT
    with contextlib.redirect_stdout(io.StringIO()):
        model = HookedViT(HookedViTConfig(**CLIP_B32))
    sd = recipe_state_dict(state_dict_shapes(dict(CLIP_B32)), 1234)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    x = torch.randn(512, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    out, cache = model.run_with_cache(x.cuda())
    assert model.last_route == "fused" and len(cache) == 214
    pick = [0, 137, 300, 511]
    with torch.no_grad():
        ref_out, ref_cache = vit_forward_with_cache(sd, dict(CLIP_B32), x[pick])
    keys = ["hook_embed", "blocks.0.attn.hook_q", "blocks.3.attn.hook_pattern", "blocks.5.hook_mlp_out", "blocks.7.mlp.hook_post",
            "blocks.9.attn.hook_z", "blocks.11.hook_resid_post", "hook_post_head_pre_normalize"]
    worst = 0.0
    for name in keys:
        got = cache[name][pick].cpu()
        worst = max(worst, rel_err(got, ref_cache[name]))
        assert_close(got, ref_cache[name], 1e-4, name)
    assert_close(out[pick].cpu(), ref_out, 1e-4, "model output")
    print(f"batch 512: worst sampled-key rel err {worst:.2e}")


def test_activations_store_with_a_real_model_matches_oracle_and_mixes_halves():
    from torch.utils.data import TensorDataset
    from vit_prisma.configs.HookedViTConfig import HookedViTConfig
    from vit_prisma.models.base_vit import HookedViT
    from vit_prisma.sae.config import VisionModelSAERunnerConfig
    from vit_prisma.sae.training.activations_store import VisionActivationsStore
    cfg_v = dict(CLIP_B32, n_layers=3, d_model=128, d_head=64, n_heads=2, d_mlp=256, patch_size=16, image_size=64, n_classes=16)
    sd = recipe_state_dict(state_dict_shapes(cfg_v), 21)
    with contextlib.redirect_stdout(io.StringIO()):
        model = HookedViT(HookedViTConfig(**cfg_v))
    model.load_state_dict(sd)
    model = model.cuda().eval()
    T = 17
    images = torch.randn(48, 3, 64, 64, generator=torch.Generator().manual_seed(4))
    with contextlib.redirect_stdout(io.StringIO()):
        cfg = VisionModelSAERunnerConfig(d_in=128, expansion_factor=4, activation_fn_str="topk", activation_fn_kwargs={"k": 4}, _device="cuda",
                                         _dtype="float32", hook_point_layer=1, layer_subtype="hook_resid_post", context_size=T, store_batch_size=4,
                                         n_batches_in_buffer=6, train_batch_size=32, n_checkpoints=0, log_to_wandb=False, image_size=64,
                                         checkpoint_path="/tmp/prisma_b200_unused", num_workers=0)
    torch.manual_seed(0)
    store = VisionActivationsStore(cfg, model, TensorDataset(images, torch.zeros(48, dtype=torch.long)), num_workers=0)
    # get_activations == the oracle's cache entry for the same images (reference activations_store.py:252-296)
    name = "blocks.1.hook_resid_post"
    acts = store.get_activations(images[:8].cuda())
    with torch.no_grad():
        _, ref_cache = vit_forward_with_cache(sd, cfg_v, images[:8], names_filter=lambda n: n == name, stop_at_layer=2)
    assert tuple(acts.shape) == (8, T, 1, 128)
    assert_close(acts[:, :, 0].cpu(), ref_cache[name], 1e-4, "store.get_activations")
    # half-buffer mix (reference :445-492): after construction the storage half holds (n_batches + n_batches // 2) / 2 image
    # batches' worth of tokens; every served / stored row is a genuine token activation of some dataset image
    per_batch = cfg.store_batch_size * T
    assert store.storage_buffer.shape == ((6 + 3) * per_batch // 2, 1, 128)
    with torch.no_grad():
        _, all_cache = vit_forward_with_cache(sd, cfg_v, images, names_filter=lambda n: n == name, stop_at_layer=2)
    bank = all_cache[name].reshape(-1, 128)
    served = torch.cat([store.next_batch() for _ in range(3)]).cpu()
    assert served.shape[1:] == (1, 128)
    rows = torch.cat([served[:, 0], store.storage_buffer[:64, 0].cpu()])
    dist = (rows[:, None, :].double() - bank[None, :, :].double()).abs().amax(dim=2).min(dim=1).values     # exact, no cdist cancellation
    assert float(dist.max()) <= 1e-4 * float(bank.abs().max()), float(dist.max())
