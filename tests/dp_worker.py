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


def bf16_trainer(rank, world, dev):
    """DP_MODE=bf16_trainer: the cfg #5 class through VisionSAETrainer(p2p_group=...) -- bf16 module storage, fp32 masters living in
    the peer-visible buffers, parameters exported after every step (after the deferred W_dec all-gather has landed).  Every rank
    takes 1/N of each batch of tests/golden/sae_bf16_v.pt; the result must be the reference's fp32 trajectory rounded to bf16."""
    import contextlib
    import io
    from tests.util import load_golden
    from vit_prisma.b200.p2p import P2PGroup
    from vit_prisma.sae.config import VisionModelSAERunnerConfig
    from vit_prisma.sae.train_sae import VisionSAETrainer
    gold = load_golden("sae_bf16_v.pt")
    B, per = gold["batch"], gold["batch"] // world
    with contextlib.redirect_stdout(io.StringIO()):
        cfg = VisionModelSAERunnerConfig(d_in=gold["d_in"], expansion_factor=gold["d_sae"] // gold["d_in"], activation_fn_str="topk",
                                         activation_fn_kwargs={"k": gold["k"]}, _device=str(dev), _dtype="bfloat16",
                                         normalize_activations=gold["norm"], b_dec_init_method="zeros", lr=gold["lr"],
                                         lr_warm_up_steps=gold["warm_up_steps"], train_batch_size=per, max_grad_norm=1.0,
                                         initialization_method="independent", log_to_wandb=False, n_checkpoints=0,
                                         checkpoint_path="/tmp/prisma_b200_unused",
                                         num_epochs=(gold["total_steps"] + 0.5) / 1_300_000, context_size=per)
        trainer = VisionSAETrainer(cfg, model=None, dataset=None, activations_store=object(), p2p_group=P2PGroup(rank, world, dev))
    assert cfg.total_training_steps == gold["total_steps"]
    sae = trainer.sparse_coder
    sae.load_state_dict({k: v.to(dev) for k, v in gold["init"].items()})
    act_freq, since_fired, n_frac, opt, sched = trainer.initialize_training_variables()
    trainer.enable_data_parallel_if_requested()
    data = gold["data"].to(dev)
    ok, worst = True, 0.0
    for s, rec32 in enumerate(gold["steps_fp32"]):
        x = data[s * B + rank * per: s * B + (rank + 1) * per].unsqueeze(1)
        _, mse, _, _, act_freq, since_fired, n_frac = trainer.train_step(sae, opt, sched, act_freq, since_fired, n_frac, x, s, s * B)
        m = mse.detach().float().reshape(1).clone()
        dist.all_reduce(m)                                   # shares of the global mean add up
        if abs(m.item() - rec32["mse"]) > 1e-4 * rec32["mse"]:
            ok = False
            print(f"[rank {rank}] step {s}: mse {m.item()} vs {rec32['mse']}", flush=True)
        sd = sae.state_dict()
        for name, ref in rec32["params_after"].items():
            if name == "W_dec":
                ref = ref / ref.norm(dim=1, keepdim=True)
            mine = sd[name]
            e = float((mine.float().cpu() - ref).abs().max()) / float(ref.abs().max())
            worst = max(worst, e)
            if mine.dtype != torch.bfloat16 or e > 1.01 * 2.0 ** -8:
                ok = False
                print(f"[rank {rank}] step {s}: {name} {mine.dtype} {e:.2e} from the fp32 trajectory", flush=True)
    from vit_prisma.b200.p2p import SaeDPEngine
    eng = sae.step_engine()
    if not isinstance(eng, SaeDPEngine):                     # the trainer must keep the peer-memory engine (round-1 advisor finding)
        ok = False
        print(f"[rank {rank}] trainer runs on {type(eng).__name__}, not SaeDPEngine", flush=True)
    # identical parameters on every rank: compare an order-sensitive checksum of the fp32 masters with rank 0's
    eng.wait_parameters()
    sums = torch.stack([(t.double() * torch.arange(1, t.numel() + 1, device=dev, dtype=torch.float64).reshape(t.shape)).sum()
                        for t in (eng.W_encT, eng.W_dec, eng.b_enc, eng.b_dec)])
    ref = sums.clone()
    dist.broadcast(ref, src=0)
    if not torch.equal(sums, ref):
        ok = False
        print(f"[rank {rank}] parameters differ from rank 0's: {sums.tolist()} vs {ref.tolist()}", flush=True)
    return ok, worst, eng.describe_exchange()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)          # handle exchange + test barriers only
    if os.environ.get("DP_MODE") == "bf16_trainer":
        ok, worst, exch = bf16_trainer(rank, world, dev)
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"DP_RESULT world={world} ok={bool(flag.item())} worst_param_rel_err={worst:.2e} (bf16 trainer) exchange: {exch}", flush=True)
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if flag.item() == 1.0 else 1)
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
