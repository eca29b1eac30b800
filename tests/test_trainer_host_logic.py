"""Host-side logic of the SAE trainer / modules that needs no GPU: variant dispatch guards and the no-CPU-fallback rule."""
import pytest
import torch

from vit_prisma.b200._lib import PrismaB200Error
from vit_prisma.sae.config import VisionModelSAERunnerConfig
from vit_prisma.sae.sae import GatedSparseAutoencoder, StandardSparseAutoencoder
from vit_prisma.sae.train_sae import VisionSAETrainer


def _cfg(**kw):
    base = dict(d_in=16, expansion_factor=2, activation_fn_str="topk", activation_fn_kwargs={"k": 4}, _device="cpu", n_checkpoints=0,
                log_to_wandb=False, checkpoint_path="/tmp/prisma_b200_unused")
    base.update(kw)
    return VisionModelSAERunnerConfig(**base)


class _Store:
    def next_batch(self):
        raise RuntimeError("not used")


class _Group:
    world, rank = 2, 0


@pytest.mark.parametrize("kw", [dict(activation_fn_str="relu", activation_fn_kwargs={}), dict(use_ghost_grads=True),
                                dict(activation_fn_str="relu", activation_fn_kwargs={}, architecture="gated")])
def test_data_parallel_is_refused_for_the_dense_ghost_and_gated_steps(kw):
    trainer = VisionSAETrainer(_cfg(**kw), model=None, dataset=None, activations_store=_Store())
    trainer.p2p_group = _Group()
    with pytest.raises(NotImplementedError, match="TopK step"):
        trainer.enable_data_parallel_if_requested()


def test_gated_module_parameters_and_cpu_refusal():
    sae = GatedSparseAutoencoder(_cfg(activation_fn_str="relu", activation_fn_kwargs={}, architecture="gated"))
    assert {n: tuple(p.shape) for n, p in sae.named_parameters()} == {
        "W_enc": (16, 32), "b_gate": (32,), "r_mag": (32,), "b_mag": (32,), "W_dec": (32, 16), "b_enc": (32,), "b_dec": (16,)}
    assert sae.W_enc.data.t().is_contiguous()                      # feature-major storage behind the reference-shaped parameter
    with pytest.raises(PrismaB200Error, match="no CPU fallback"):
        sae(torch.randn(4, 16))
    with pytest.raises(AssertionError):
        GatedSparseAutoencoder(_cfg(activation_fn_str="relu", activation_fn_kwargs={}, architecture="gated", use_ghost_grads=True))


def test_step_engine_rejects_unbuilt_training_activations():
    sae = StandardSparseAutoencoder(_cfg(activation_fn_str="tanh-relu", activation_fn_kwargs={}))
    with pytest.raises(NotImplementedError, match="tanh-relu"):
        sae.step_engine()
    sae = StandardSparseAutoencoder(_cfg(activation_fn_str="relu", activation_fn_kwargs={}, lp_norm=2))
    with pytest.raises(NotImplementedError, match="lp_norm"):
        sae.step_engine()


def test_bf16_module_keeps_reference_storage_and_refuses_cpu_compute(tmp_path):
    """cfg #5 class: `_dtype="bfloat16"` (the reference's dtype table has no such entry) gives bf16 parameters / state dict with the
    reference's names and shapes; the fp32 masters the step engine trains do not exist until a CUDA engine is bound; on a host-
    resident module every compute entry (forward, step_engine) fails loudly instead of computing on the CPU."""
    cfg = _cfg(_dtype="bfloat16")
    assert cfg.dtype == torch.bfloat16
    sae = StandardSparseAutoencoder(cfg)
    assert sae.low_precision and sae._masters is None
    sd = sae.state_dict()
    assert {k: (tuple(v.shape), v.dtype) for k, v in sd.items()} == {
        "W_dec": ((32, 16), torch.bfloat16), "W_enc": ((16, 32), torch.bfloat16), "b_enc": ((32,), torch.bfloat16), "b_dec": ((16,), torch.bfloat16)}
    assert sae.W_enc.data.t().is_contiguous()
    sae.export_masters()                                           # nothing to export yet: a no-op, not an error
    with pytest.raises(PrismaB200Error, match="no CPU fallback"):
        sae(torch.randn(4, 16, dtype=torch.bfloat16))
    with pytest.raises(PrismaB200Error, match="no CPU fallback"):
        sae.step_engine()
    # config round trip keeps the dtype string; fp32 modules are not "low precision"
    path = tmp_path / "config.json"
    cfg.save_config(str(path))
    assert VisionModelSAERunnerConfig.load_config(str(path)).dtype == torch.bfloat16
    assert not StandardSparseAutoencoder(_cfg()).low_precision
