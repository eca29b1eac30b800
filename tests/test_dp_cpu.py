"""world_size-2 gloo test (CPU): the data-parallel decomposition the P2P kernels implement is exact.

Each rank holds half of a batch and computes, with the ORACLE, its local gradients using the GLOBAL column mean of x
(all-reduced d floats) and the GLOBAL token count in the mean-loss factor; the SUM of the two ranks' gradients, dead-feature
counts and squared-norm partials must equal the single-process values on the full batch (sae/train_sae.py:278-411 semantics).
Also covers the host-side pieces of vit_prisma/b200/p2p.py that need no GPU: shard bounds and the handle exchange."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.sae_oracle import sae_forward, sae_grads
from tests.util import load_golden


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vit_prisma.b200.p2p import P2PGroup, shard_bounds
        gold = load_golden("sae_tiny_a.pt")
        g = torch.Generator().manual_seed(gold["data_seed"])
        B, d, k, F = gold["batch"], gold["d_in"], gold["k"], gold["d_sae"]
        x = (torch.randn(B * gold["n_steps"], d, generator=g) * 2.0 + torch.randn(d, generator=g))[:B]
        p = {n: v.clone() for n, v in gold["init"].items()}
        p["W_dec"] /= p["W_dec"].norm(dim=1, keepdim=True)
        per = B // world
        xs = x[rank * per:(rank + 1) * per]
        xsum = xs.sum(0)
        dist.all_reduce(xsum)                                   # what pb_p2p_sum_xsum does with peer loads
        fwd = sae_forward(p, xs, k, xbar=(xsum / B).unsqueeze(0), global_rows=B)
        grads = sae_grads(p, xs, fwd, global_rows=B)
        fired = (fwd["feature_acts"] > 0).float().sum(0)
        mse_share = fwd["mse"].clone()
        for t in list(grads.values()) + [fired, mse_share]:
            dist.all_reduce(t)                                  # reduce-scatter + all-gather == all-reduce
        # owner-computes partition of the norm: each rank squares only its feature-row slice, small vectors counted once
        f0, f1 = shard_bounds(F, rank, world)
        part = (grads["W_dec"][f0:f1] ** 2).sum() + (grads["W_enc"][:, f0:f1] ** 2).sum()
        if rank == 0:
            part = part + (grads["b_enc"] ** 2).sum() + (grads["b_dec"] ** 2).sum()
        dist.all_reduce(part)
        # handle exchange plumbing (all_gather_object) with fake 64-byte handles
        everyone = P2PGroup._exchange_dist({"W_dec": bytes([rank]) * 64})
        if rank == 0:
            full = sae_forward(p, x, k)
            ref = sae_grads(p, x, full)
            ok = all(torch.allclose(grads[n], ref[n], rtol=1e-5, atol=1e-8) for n in ref)
            ok &= torch.equal(fired, (full["feature_acts"] > 0).float().sum(0))
            ok &= abs(mse_share.item() - full["mse"].item()) < 1e-6 * full["mse"].item()
            ref_norm = sum((v.double() ** 2).sum() for v in ref.values())
            ok &= abs(part.item() - ref_norm.item()) < 1e-5 * ref_norm.item()
            ok &= [e["W_dec"][0] for e in everyone] == [0, 1]
            ret["ok"] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_data_parallel_decomposition_is_exact():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok") is True


def test_shard_bounds():
    from vit_prisma.b200.p2p import shard_bounds
    assert [shard_bounds(24576, r, 8) for r in (0, 7)] == [(0, 3072), (21504, 24576)]
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 3)
