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
