"""save_model -> load_from_pretrained round trips (reference sae/sae.py:299-523): class chosen from the loaded config, the
config.json fallback for weights-only files, mapping overrides through ``current_cfg``.  CPU only (construction + state dicts)."""
import torch

from vit_prisma.sae.config import VisionModelSAERunnerConfig
from vit_prisma.sae.sae import GatedSparseAutoencoder, SparseAutoencoder, StandardSparseAutoencoder


def _cfg(**kw):
    base = dict(d_in=16, expansion_factor=4, _device="cpu", _dtype="float32", n_checkpoints=0, log_to_wandb=False,
                activation_fn_str="topk", activation_fn_kwargs={"k": 4}, checkpoint_path="/tmp/prisma_b200_unused")
    base.update(kw)
    return VisionModelSAERunnerConfig(**base)


def _same_state(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]), k


def test_round_trip_through_the_abstract_base_picks_the_architecture(tmp_path):
    std = StandardSparseAutoencoder(_cfg())
    gated = GatedSparseAutoencoder(_cfg(architecture="gated", activation_fn_str="relu", activation_fn_kwargs={}))
    for name, model, cls in (("std.pt", std, StandardSparseAutoencoder), ("gated.pkl.gz", gated, GatedSparseAutoencoder)):
        path = str(tmp_path / name)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.01)
        model.save_model(path)
        loaded = SparseAutoencoder.load_from_pretrained(path)            # the reference's canonical call (evals.py:170, load_model.py:64)
        assert type(loaded) is cls
        _same_state(model, loaded)
        assert loaded.W_enc.shape == (16, 64) and loaded.W_dec.shape == (64, 16)


def test_weights_only_file_uses_config_json_next_to_it_and_mapping_overrides(tmp_path):
    model = StandardSparseAutoencoder(_cfg())
    model.cfg.save_config(str(tmp_path / "config.json"))
    torch.save({k: v.contiguous() for k, v in model.state_dict().items()}, tmp_path / "weights.pt")
    loaded = SparseAutoencoder.load_from_pretrained(str(tmp_path / "weights.pt"), current_cfg={"l1_coefficient": 0.5, "not_a_field": 1})
    _same_state(model, loaded)
    assert loaded.cfg.l1_coefficient == 0.5 and not hasattr(loaded.cfg, "not_a_field")
    # an explicit config path wins over an embedded cfg
    model.save_model(str(tmp_path / "combined.pt"))
    again = StandardSparseAutoencoder.load_from_pretrained(str(tmp_path / "combined.pt"), config_path=str(tmp_path / "config.json"))
    _same_state(model, again)


def test_missing_config_is_reported(tmp_path):
    import pytest
    torch.save({"W_dec": torch.zeros(2, 2)}, tmp_path / "w.pt")
    with pytest.raises(FileNotFoundError):
        SparseAutoencoder.load_from_pretrained(str(tmp_path / "w.pt"))
    with pytest.raises(FileNotFoundError):
        SparseAutoencoder.load_from_pretrained(str(tmp_path / "nope.pt"))
