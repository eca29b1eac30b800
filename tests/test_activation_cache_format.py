"""Disk activation cache (reference sae/training/activations_store.py:21-152, 371-415, 505-574): fp16 ``{idx}.pt``
shards of ``[tokens, n_layers, d_in]`` -- writer, buffer loader and the cache-backed store, on CPU with a stand-in model."""
import os

import pytest
import torch
from torch.utils.data import TensorDataset

from vit_prisma.sae.config import VisionModelSAERunnerConfig
from vit_prisma.sae.training.activations_store import CacheVisionActivationStore, VisionActivationsStore

T, D = 5, 8


class _FakeViT:
    """run_with_cache stand-in: the 'activation' of token t of an image is a fixed linear function of the image mean."""

    def to(self, *a, **k):
        return self

    def run_with_cache(self, images, names_filter=None, stop_at_layer=None):
        m = images.float().mean(dim=(1, 2, 3))                                # [b]
        acts = m[:, None, None] * torch.arange(1, T + 1)[None, :, None] + torch.arange(D)[None, None, :] * 0.125
        return None, {name: acts.clone() for name in names_filter}


def _cfg(path, **kw):
    base = dict(d_in=D, expansion_factor=2, activation_fn_str="topk", activation_fn_kwargs={"k": 2}, _device="cpu", _dtype="float32",
                hook_point_layer=1, layer_subtype="hook_resid_post", context_size=T, store_batch_size=4, train_batch_size=6,
                cached_activations_path=str(path), n_checkpoints=0, log_to_wandb=False, checkpoint_path="/tmp/prisma_b200_unused", num_workers=0)
    base.update(kw)
    return VisionModelSAERunnerConfig(**base)


def test_writer_loader_and_cache_store_round_trip(tmp_path):
    images = torch.arange(10, dtype=torch.float32)[:, None, None, None].expand(10, 3, 4, 4).contiguous()
    data = TensorDataset(images, torch.zeros(10, dtype=torch.long))
    cfg = _cfg(tmp_path)
    store = VisionActivationsStore(cfg, _FakeViT(), data, create_dataloader=False)
    n_files = store.generate_cached_activations_from_dataset(tokens_per_file=16)
    assert n_files == 4 and sorted(os.listdir(tmp_path)) == ["0.pt", "1.pt", "2.pt", "3.pt"]
    shards = [torch.load(tmp_path / f"{i}.pt", weights_only=True) for i in range(4)]
    assert [tuple(s.shape) for s in shards] == [(16, 1, D), (16, 1, D), (16, 1, D), (2, 1, D)]
    assert all(s.dtype == torch.float16 for s in shards)
    _, cache = _FakeViT().run_with_cache(images, names_filter=[cfg.hook_point])
    expect = cache[cfg.hook_point].reshape(-1, 1, D).half()
    assert torch.equal(torch.cat(shards), expect)                             # token order = (image, position), as the reference writes it
    # the buffer loader of the live store reads the same files back (reference _load_cached_activations)
    buf = store._load_cached_activations(total_size=10, context_size=T, num_layers=1, d_in=D)
    assert buf.dtype == torch.float32 and torch.equal(buf, expect.float())
    assert store._load_cached_activations(total_size=3, context_size=T, num_layers=1, d_in=D).shape == (15, 1, D)
    # the cache-backed store: reference half-buffer mixing (activations_store.py:21-152) over the shards, wrapping around
    with pytest.raises(ValueError):
        CacheVisionActivationStore(cfg)                                       # reference :36-37
    cfg.use_cached_activations = True
    cached = CacheVisionActivationStore(cfg)
    half_tokens = (cfg.n_batches_in_buffer // 2) * cfg.store_batch_size * cfg.context_size
    assert cached.storage_buffer.shape == (half_tokens, 1, D)                # kept half of a (fresh half + stored half) mix
    seen = []
    for _ in range(12):
        b = cached.next_batch()
        assert b.shape[1:] == (1, D) and 0 < b.shape[0] <= cfg.train_batch_size
        seen.append(b)
    got = torch.cat(seen)
    rows = {tuple(r.flatten().tolist()) for r in expect.float()}
    assert all(tuple(r.flatten().tolist()) in rows for r in got)
    assert cached._file_idx >= 1 or cached._file_off > 0                      # the file cursor persists across refills


def test_trainer_b_dec_init_reads_the_cache_store_buffer(tmp_path):
    """VisionSAETrainer.initialize_geometric_medians() with a cache-backed store (mean / zeros / geometric_median)."""
    from vit_prisma.sae.train_sae import VisionSAETrainer
    images = torch.arange(10, dtype=torch.float32)[:, None, None, None].expand(10, 3, 4, 4).contiguous()
    cfg = _cfg(tmp_path)
    VisionActivationsStore(cfg, _FakeViT(), TensorDataset(images, torch.zeros(10, dtype=torch.long)),
                           create_dataloader=False).generate_cached_activations_from_dataset(tokens_per_file=16)
    for method in ("zeros", "mean", "geometric_median"):
        cfg2 = _cfg(tmp_path, use_cached_activations=True, b_dec_init_method=method, activation_fn_str="topk", activation_fn_kwargs={"k": 2})
        trainer = VisionSAETrainer(cfg2, model=None, dataset=None)
        assert isinstance(trainer.activations_store, CacheVisionActivationStore)
        trainer.initialize_geometric_medians()
        b = trainer.sparse_coder.b_dec.detach()
        buf = trainer.activations_store.storage_buffer[:, 0, :].float()
        if method == "zeros":
            assert torch.count_nonzero(b) == 0
        elif method == "mean":
            assert torch.allclose(b.float().cpu(), buf.mean(0).cpu(), atol=1e-6)
        else:
            assert torch.isfinite(b).all() and (b.float().cpu() - buf.mean(0).cpu()).abs().max() < buf.abs().max()


def test_patches_only_drops_the_class_token(tmp_path):
    images = torch.ones(4, 3, 4, 4)
    cfg = _cfg(tmp_path, context_size=T - 1, use_patches_only=True)
    store = VisionActivationsStore(cfg, _FakeViT(), TensorDataset(images, torch.zeros(4, dtype=torch.long)), create_dataloader=False)
    assert store.generate_cached_activations_from_dataset(tokens_per_file=1000) == 1
    shard = torch.load(tmp_path / "0.pt", weights_only=True)
    assert shard.shape == (4 * (T - 1), 1, D) and float(shard[0, 0, 0]) == 2.0   # first kept token is position 1 (value 1 * 2)
