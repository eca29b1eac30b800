"""Gated-SAE fixture from the UNMODIFIED reference module (run in the build container only; see make_golden.py)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def synthetic_acts(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g) * 2.0 + torch.randn(d, generator=g)


def make():
    from vit_prisma.sae.config import VisionModelSAERunnerConfig
    from vit_prisma.sae.sae import GatedSparseAutoencoder
    from vit_prisma.sae.training.get_scheduler import get_scheduler

    for tag, d_in, expansion, batch, norm, l1c in (("g", 32, 8, 64, "layer_norm", 4e-3), ("h", 48, 4, 80, "none", 1e-2)):
        torch.manual_seed(0)
        cfg = VisionModelSAERunnerConfig(d_in=d_in, expansion_factor=expansion, activation_fn_str="relu", architecture="gated",
                                         l1_coefficient=l1c, _device="cpu", _dtype="float32", normalize_activations=norm,
                                         b_dec_init_method="mean", lr=1e-3, lr_warm_up_steps=3, train_batch_size=batch, max_grad_norm=1.0,
                                         log_to_wandb=False, n_checkpoints=0, checkpoint_path="/tmp/unused", use_ghost_grads=False)
        sae = GatedSparseAutoencoder(cfg)
        n_steps, total_steps = 6, 40
        data = synthetic_acts(batch * n_steps, d_in, seed=11)
        sae.initialize_b_dec_with_mean(data)                      # train_sae.py:270-274
        with torch.no_grad():                                     # zeros at init would leave r_mag / b_mag / b_gate untested
            g = torch.Generator().manual_seed(1)
            sae.r_mag.copy_(0.1 * torch.randn(cfg.d_sae, generator=g))
            sae.b_mag.copy_(0.05 * torch.randn(cfg.d_sae, generator=g))
            sae.b_gate.copy_(0.05 * torch.randn(cfg.d_sae, generator=g))
        init = {k_: v.detach().clone() for k_, v in sae.state_dict().items()}
        opt = torch.optim.Adam(sae.parameters(), lr=cfg.lr)        # train_sae.py:229
        sched = get_scheduler(cfg.lr_scheduler_name, optimizer=opt, warm_up_steps=cfg.lr_warm_up_steps,
                              training_steps=total_steps, lr_end=cfg.lr / 10)
        since_fired, act_freq = torch.zeros(cfg.d_sae), torch.zeros(cfg.d_sae)
        steps = []
        for s in range(n_steps):
            x = data[s * batch:(s + 1) * batch]
            lr_now = opt.param_groups[0]["lr"]
            sae.train()
            sae.set_decoder_norm_to_unit_norm()                   # :306-307
            opt.zero_grad()
            sae_out, feature_acts, loss, mse, l1, ghost_loss, aux = sae(x, (since_fired > cfg.dead_feature_window).bool())
            with torch.no_grad():                                 # :356-365
                did_fire = (feature_acts > 0).float().sum(-2) > 0
                since_fired += 1
                since_fired[did_fire] = 0
                act_freq += (feature_acts.abs() > 0).float().sum(0)
                l0 = (feature_acts > 0).float().sum(-1).mean()
            loss.backward()
            raw = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in sae.named_parameters()}
            gnorm = torch.nn.utils.clip_grad_norm_(sae.parameters(), max_norm=cfg.max_grad_norm)
            sae.remove_gradient_parallel_to_decoder_directions()
            final = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in sae.named_parameters()}
            opt.step()
            sched.step()
            rec = dict(lr=lr_now, loss=loss.item(), mse=mse.item(), l1=float(l1), aux=float(aux), l0=l0.item(), grad_norm=float(gnorm),
                       sae_out=sae_out.detach().clone(), feature_acts=feature_acts.detach().clone())
            if s in (0, 3):
                rec["raw_grads"], rec["final_grads"] = raw, final
            if s in (0, 2, 5):
                rec["params_after"] = {k_: v.detach().clone() for k_, v in sae.state_dict().items()}
            steps.append(rec)
        path = os.path.join(HERE, f"sae_gated_{tag}.pt")
        torch.save(dict(d_in=d_in, d_sae=cfg.d_sae, batch=batch, norm=norm, l1_coefficient=l1c, lr=cfg.lr, warm_up_steps=cfg.lr_warm_up_steps,
                        total_steps=total_steps, lr_end=cfg.lr / 10, data_seed=11, n_steps=n_steps, init=init, steps=steps,
                        since_fired=since_fired.clone(), act_freq=act_freq.clone()), path)
        print("wrote", path, os.path.getsize(path), "bytes; loss", [round(r["loss"], 4) for r in steps], "aux", [round(r["aux"], 3) for r in steps],
              "b_enc grad None:", raw["b_enc"] is None)


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    import _ref_shims
    _ref_shims.install()
    make()
