"""torchrun worker for the multi-GPU SAE test (tests/test_sae_dp_gpu.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dp_worker.py

Every rank takes its 1/N slice of each golden batch and trains with SaeDPEngine (NVLink reduce-scatter / sharded Adam /
all-gather); the result must equal the reference's single-process training on the full batches (tests/golden/sae_tiny_b.pt),
because the data-parallel step keeps single-GPU semantics (global batch mean, global clip norm, summed counters)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vit-prisma_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)          # handle exchange + test barriers only
    from oracle.sae_oracle import lr_multiplier
    from tests.util import load_golden, rel_err
    from vit_prisma.b200.p2p import P2PGroup, SaeDPEngine
    from vit_prisma.b200.sae_engine import unit_norm_rows_

    gold = load_golden(os.environ.get("DP_GOLDEN", "sae_tiny_b.pt"))
    g = torch.Generator().manual_seed(gold["data_seed"])
    B, d, k, F = gold["batch"], gold["d_in"], gold["k"], gold["d_sae"]
    data = torch.randn(B * gold["n_steps"], d, generator=g) * 2.0 + torch.randn(d, generator=g)
    assert B % world == 0 and F % world == 0
    init = gold["init"]
    group = P2PGroup(rank, world, dev)
    eng = SaeDPEngine(group, init["W_enc"].t().contiguous().to(dev), init["W_dec"].clone().to(dev), init["b_enc"].clone().to(dev),
                      init["b_dec"].clone().to(dev), k=k, normalize_activations=gold["norm"], max_grad_norm=1.0)
    unit_norm_rows_(eng.W_dec)
    eng.refresh_lo()
    since_fired = torch.zeros(F, device=dev)
    act_freq = torch.zeros(F, device=dev)
    per = B // world
    ok, worst = True, 0.0
    for s, rec in enumerate(gold["steps"]):
        x = data[s * B + rank * per: s * B + (rank + 1) * per].to(dev)
        lr = gold["lr"] * lr_multiplier(s, gold["warm_up_steps"], gold["total_steps"], gold["lr_end"])
        eng.train_step(x, lr, since_fired=since_fired, act_freq=act_freq)
        torch.cuda.synchronize()
        sc = eng.scalars_dict()
        mse = torch.tensor([sc["mse"]], device=dev)
        dist.all_reduce(mse)                                 # shares of the global mean add up
        if abs(mse.item() - rec["mse"]) > 1e-4 * abs(rec["mse"]) or abs(sc["grad_norm"] - rec["grad_norm"]) > 1e-4 * rec["grad_norm"]:
            ok = False
            print(f"[rank {rank}] step {s}: mse {mse.item()} vs {rec['mse']}, grad_norm {sc['grad_norm']} vs {rec['grad_norm']}", flush=True)
        ref_idx = rec["topk_idx"][rank * per:(rank + 1) * per]
        if not torch.equal(eng.idx.cpu().long(), ref_idx):
            ok = False
            print(f"[rank {rank}] step {s}: TopK indices differ", flush=True)
        if "params_after" in rec:
            ref = rec["params_after"]
            ref_dec = ref["W_dec"] / ref["W_dec"].norm(dim=1, keepdim=True)
            for name, got, want in (("W_dec", eng.W_dec, ref_dec), ("W_enc", eng.W_encT.t(), ref["W_enc"]), ("b_dec", eng.b_dec, ref["b_dec"])):
                e = rel_err(got.cpu(), want)
                worst = max(worst, e)
                if e > 1e-4:
                    ok = False
                    print(f"[rank {rank}] step {s}: {name} rel err {e:.2e}", flush=True)
    if not (torch.equal(since_fired.cpu(), gold["since_fired"]) and torch.equal(act_freq.cpu(), gold["act_freq"])):
        ok = False
        print(f"[rank {rank}] dead-feature counters differ", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"DP_RESULT world={world} ok={bool(flag.item())} worst_param_rel_err={worst:.2e} exchange: {eng.describe_exchange()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
