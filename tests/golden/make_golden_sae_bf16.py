"""bf16 SAE fixture (cfg #5 shape class) from the UNMODIFIED reference modules -- run in the build container only.

The reference's ``dtype_mapping`` (sae/config.py:14-45) has no "bfloat16" entry, so ``_dtype="bfloat16"`` raises KeyError there.
SHIM (stated): one entry is added to that dict at run time -- nothing else of the reference is touched; parameters, forward,
autograd, ``torch.optim.Adam`` (moments in the parameter dtype) and the scheduler then run in bf16 exactly as its code does.

Two trajectories from the same bf16-rounded initial parameters and the same bf16-rounded data:
  * ``steps``      -- the reference in bfloat16 (what cfg #5 asks for);
  * ``steps_fp32`` -- the same reference code in float32: the exact-arithmetic trajectory both bf16 runs approximate.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def synthetic_acts(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g) * 2.0 + torch.randn(d, generator=g)


def run(dtype_name, init, data, d_in, expansion, k, batch, n_steps, total_steps):
    from vit_prisma.sae.config import VisionModelSAERunnerConfig
    from vit_prisma.sae.sae import StandardSparseAutoencoder
    from vit_prisma.sae.training.get_scheduler import get_scheduler
    torch.manual_seed(0)
    cfg = VisionModelSAERunnerConfig(d_in=d_in, expansion_factor=expansion, activation_fn_str="topk", activation_fn_kwargs={"k": k},
                                     _device="cpu", _dtype=dtype_name, normalize_activations="layer_norm", b_dec_init_method="mean",
                                     lr=1e-3, lr_warm_up_steps=2, train_batch_size=batch, max_grad_norm=1.0,
                                     initialization_method="independent", log_to_wandb=False, n_checkpoints=0,
                                     checkpoint_path="/tmp/unused", use_ghost_grads=False)
    sae = StandardSparseAutoencoder(cfg)
    dt = cfg.dtype
    if init is None:
        sae.initialize_b_dec_with_mean(data.to(dt))
        init = {k_: v.detach().clone() for k_, v in sae.state_dict().items()}
    else:
        sae.load_state_dict({k_: v.to(dt) for k_, v in init.items()})
    opt = torch.optim.Adam(sae.parameters(), lr=cfg.lr)
    sched = get_scheduler(cfg.lr_scheduler_name, optimizer=opt, warm_up_steps=cfg.lr_warm_up_steps, training_steps=total_steps, lr_end=cfg.lr / 10)
    steps = []
    for s in range(n_steps):
        x = data[s * batch:(s + 1) * batch].to(dt)
        lr_now = opt.param_groups[0]["lr"]
        sae.train()
        sae.set_decoder_norm_to_unit_norm()
        opt.zero_grad()
        sae_out, feature_acts, loss, mse, l1, ghost_loss, aux = sae(x, None)
        with torch.no_grad():
            l0 = (feature_acts > 0).float().sum(-1).mean()
        loss.backward()
        gnorm = torch.nn.utils.clip_grad_norm_(sae.parameters(), max_norm=cfg.max_grad_norm)
        sae.remove_gradient_parallel_to_decoder_directions()
        opt.step()
        sched.step()
        steps.append(dict(lr=lr_now, loss=float(loss), mse=float(mse), l0=float(l0), grad_norm=float(gnorm), sae_out=sae_out.detach().clone(),
                          params_after={k_: v.detach().clone() for k_, v in sae.state_dict().items()}))
    return init, steps, cfg


def main():
    import vit_prisma.sae.config as ref_cfg
    ref_cfg.dtype_mapping["bfloat16"] = torch.bfloat16             # the stated shim
    d_in, expansion, k, batch, n_steps, total_steps = 64, 8, 16, 96, 4, 40
    data = synthetic_acts(batch * n_steps, d_in, seed=11).to(torch.bfloat16)       # bf16-representable in both runs
    init, steps, cfg = run("bfloat16", None, data, d_in, expansion, k, batch, n_steps, total_steps)
    _, steps32, _ = run("float32", {k_: v.float() for k_, v in init.items()}, data.float(), d_in, expansion, k, batch, n_steps, total_steps)
    path = os.path.join(HERE, "sae_bf16_v.pt")
    torch.save(dict(d_in=d_in, d_sae=cfg.d_sae, k=k, batch=batch, norm="layer_norm", lr=cfg.lr, warm_up_steps=cfg.lr_warm_up_steps,
                    total_steps=total_steps, lr_end=cfg.lr / 10, n_steps=n_steps, init=init, data=data, steps=steps, steps_fp32=steps32), path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for a, b in zip(steps, steps32):
        dw = (a["params_after"]["W_dec"].float() - b["params_after"]["W_dec"]).abs().max().item()
        print(f"  bf16 mse {a['mse']:.5f} / fp32 mse {b['mse']:.5f}   grad_norm {a['grad_norm']:.4f} / {b['grad_norm']:.4f}   max |W_dec bf16 - fp32| {dw:.2e}")


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    import _ref_shims
    _ref_shims.install()
    main()
