"""SAE fixtures from the UNMODIFIED reference modules (run in the build container only; see make_golden.py)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def synthetic_acts(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g) * 2.0 + torch.randn(d, generator=g)      # non-zero mean


def make_sae():
    from vit_prisma.sae.config import VisionModelSAERunnerConfig
    from vit_prisma.sae.sae import StandardSparseAutoencoder
    from vit_prisma.sae.training.get_scheduler import get_scheduler

    # tag, d_in, expansion, k, batch, norm, activation, l1 coefficient, ghost grads, dead_feature_window
    for tag, d_in, expansion, k, batch, norm, act, l1c, ghost, dead_window in (
            ("a", 32, 8, 8, 64, "layer_norm", "topk", 2e-4, False, 5000), ("b", 64, 8, 16, 96, "layer_norm", "topk", 2e-4, False, 5000),
            ("c", 32, 8, 4, 48, "none", "topk", 2e-4, False, 5000),
            ("d", 32, 8, 0, 64, "layer_norm", "relu", 4e-3, False, 5000),     # dense ReLU + L1 (the reference's default activation)
            ("e", 32, 8, 4, 64, "layer_norm", "topk", 2e-4, True, 1),         # ghost grads on features dead for > 1 step
            ("f", 48, 8, 0, 80, "none", "relu", 4e-3, True, 1)):              # ghost grads on the dense path
        torch.manual_seed(0)
        cfg = VisionModelSAERunnerConfig(d_in=d_in, expansion_factor=expansion, activation_fn_str=act,
                                         activation_fn_kwargs=({"k": k} if act == "topk" else {}), l1_coefficient=l1c,
                                         _device="cpu", _dtype="float32", normalize_activations=norm, b_dec_init_method="mean",
                                         lr=1e-3, lr_warm_up_steps=3, train_batch_size=batch, max_grad_norm=1.0,
                                         initialization_method="independent", log_to_wandb=False, n_checkpoints=0,
                                         checkpoint_path="/tmp/unused", use_ghost_grads=ghost, dead_feature_window=dead_window)
        sae = StandardSparseAutoencoder(cfg)
        n_steps, total_steps = 6, 40
        data = synthetic_acts(batch * n_steps, d_in, seed=7)
        sae.initialize_b_dec_with_mean(data)                      # train_sae.py:270-274
        if tag == "f":                                            # a ReLU dictionary has no dead features at init: silence every 5th one
            with torch.no_grad():
                sae.b_enc[::5] = -30.0
        init = {k_: v.detach().clone() for k_, v in sae.state_dict().items()}
        opt = torch.optim.Adam(sae.parameters(), lr=cfg.lr)        # train_sae.py:229
        sched = get_scheduler(cfg.lr_scheduler_name, optimizer=opt, warm_up_steps=cfg.lr_warm_up_steps,
                              training_steps=total_steps, lr_end=cfg.lr / 10)   # :230-236
        since_fired = torch.zeros(cfg.d_sae)
        act_freq = torch.zeros(cfg.d_sae)
        steps = []
        for s in range(n_steps):
            x = data[s * batch:(s + 1) * batch]
            lr_now = opt.param_groups[0]["lr"]
            sae.train()
            sae.set_decoder_norm_to_unit_norm()                   # :306-307
            opt.zero_grad()
            mask_used = (since_fired > cfg.dead_feature_window).bool()   # :330-332 ghost_grad_neuron_mask
            sae_out, feature_acts, loss, mse, l1, ghost_loss, aux = sae(x, mask_used)   # :346-354
            with torch.no_grad():                                 # :356-365
                did_fire = (feature_acts > 0).float().sum(-2) > 0
                since_fired += 1
                since_fired[did_fire] = 0
                act_freq += (feature_acts.abs() > 0).float().sum(0)
                l0 = (feature_acts > 0).float().sum(-1).mean()
                _, _, hidden_pre = sae.encode(x, return_hidden_pre=True)
                top = torch.topk(hidden_pre, k=k if act == "topk" else 4, dim=-1)
            loss.backward()                                       # :392
            raw = {n: p.grad.detach().clone() for n, p in sae.named_parameters()}
            gnorm = torch.nn.utils.clip_grad_norm_(sae.parameters(), max_norm=cfg.max_grad_norm)   # :394-397
            sae.remove_gradient_parallel_to_decoder_directions()   # :399
            final = {n: p.grad.detach().clone() for n, p in sae.named_parameters()}
            opt.step()
            sched.step()                                          # :400-401
            rec = dict(lr=lr_now, loss=loss.item(), mse=mse.item(), l0=l0.item(), grad_norm=float(gnorm), topk_idx=top.indices.clone(),
                       topk_val=top.values.clone(), sae_out=sae_out.detach().clone(), l1=(None if l1 is None else float(l1)),
                       ghost=float(ghost_loss), n_dead=int(mask_used.sum()))
            if s == 0 or (ghost and s in (3, 5)):
                rec["raw_grads"], rec["final_grads"] = raw, final
            if s in (0, 2, 5):
                rec["params_after"] = {k_: v.detach().clone() for k_, v in sae.state_dict().items()}
            steps.append(rec)
        path = os.path.join(HERE, f"sae_tiny_{tag}.pt")
        torch.save(dict(act=act, l1_coefficient=l1c, use_ghost_grads=ghost, dead_feature_window=dead_window, lp_norm=cfg.lp_norm,
                        d_in=d_in, d_sae=cfg.d_sae, k=k, batch=batch, norm=norm, lr=cfg.lr, warm_up_steps=cfg.lr_warm_up_steps,
                        total_steps=total_steps, lr_end=cfg.lr / 10, data_seed=7, n_steps=n_steps, init=init, steps=steps,
                        since_fired=since_fired.clone(), act_freq=act_freq.clone()), path)
        print("wrote", path, os.path.getsize(path), "bytes", "n_dead per step", [r["n_dead"] for r in steps], "ghost", [round(r["ghost"], 5) for r in steps])


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    import _ref_shims
    _ref_shims.install()
    make_sae()
