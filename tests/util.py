"""Parity helpers shared by the tests."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    """max |got - ref| relative to max |ref| of the tensor (the metric north_star's 1e-4 / 1e-2 refer to)."""
    got = got.detach().to("cpu", torch.float64)
    ref = ref.detach().to("cpu", torch.float64)
    denom = max(ref.abs().max().item(), 1e-30)
    return (got - ref).abs().max().item() / denom


def assert_close(got, ref, tol, what=""):
    assert tuple(got.shape) == tuple(ref.shape), f"{what}: shape {tuple(got.shape)} != {tuple(ref.shape)}"
    assert got.dtype == ref.dtype, f"{what}: dtype {got.dtype} != {ref.dtype}"
    e = rel_err(got, ref)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"
    return e
