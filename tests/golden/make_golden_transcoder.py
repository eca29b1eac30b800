"""Transcoder fixtures from the UNMODIFIED reference module (sae/transcoder.py:6-116 driven as train_sae.py:278-411 drives it);
run in the build container only (see make_golden.py).

    python tests/golden/make_golden_transcoder.py   ->  tests/golden/transcoder_{t,u}.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def synthetic_pair(n, d, seed):
    """(input, target) activations: the target is a fixed random linear map of the input plus noise (an MLP-out-like relation)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g) * 2.0 + torch.randn(d, generator=g)
    M = torch.randn(d, d, generator=g) / d ** 0.5
    y = torch.tanh(x @ M) * 1.5 + 0.3 * torch.randn(n, d, generator=g) + torch.randn(d, generator=g)
    return x, y


def make():
    from vit_prisma.sae.config import VisionModelSAERunnerConfig
    from vit_prisma.sae.transcoder import Transcoder
    from vit_prisma.sae.training.get_scheduler import get_scheduler

    # tag, d, expansion, batch, norm, activation, k, l1, skip connection
    for tag, d, expansion, batch, norm, act, k, l1c, skip in (("t", 32, 8, 64, "layer_norm", "relu", 0, 4e-3, True),
                                                              ("u", 48, 8, 80, "none", "topk", 8, 2e-4, False)):
        torch.manual_seed(0)
        cfg = VisionModelSAERunnerConfig(d_in=d, d_out=d, expansion_factor=expansion, activation_fn_str=act,
                                         activation_fn_kwargs=({"k": k} if act == "topk" else {}), l1_coefficient=l1c, is_transcoder=True,
                                         transcoder_with_skip_connection=skip, _device="cpu", _dtype="float32", normalize_activations=norm,
                                         b_dec_init_method="mean", lr=1e-3, lr_warm_up_steps=3, train_batch_size=batch, max_grad_norm=1.0,
                                         log_to_wandb=False, n_checkpoints=0, checkpoint_path="/tmp/unused", use_ghost_grads=False)
        tc = Transcoder(cfg)
        n_steps, total_steps = 5, 40
        x_all, y_all = synthetic_pair(batch * n_steps, d, seed=13)
        with torch.no_grad():                                     # non-trivial biases (both start at zero in the reference)
            g = torch.Generator().manual_seed(2)
            tc.b_dec.copy_(x_all.mean(0))
            tc.b_dec_out.copy_(0.1 * torch.randn(d, generator=g))
            tc.b_enc.copy_(0.02 * torch.randn(cfg.d_sae, generator=g))
        init = {k_: v.detach().clone() for k_, v in tc.state_dict().items()}
        opt = torch.optim.Adam(tc.parameters(), lr=cfg.lr)
        sched = get_scheduler(cfg.lr_scheduler_name, optimizer=opt, warm_up_steps=cfg.lr_warm_up_steps, training_steps=total_steps,
                              lr_end=cfg.lr / 10)
        since_fired, act_freq = torch.zeros(cfg.d_sae), torch.zeros(cfg.d_sae)
        steps = []
        for s in range(n_steps):
            x, y = x_all[s * batch:(s + 1) * batch], y_all[s * batch:(s + 1) * batch]
            lr_now = opt.param_groups[0]["lr"]
            tc.train()
            tc.set_decoder_norm_to_unit_norm()                    # train_sae.py:306-307
            opt.zero_grad()
            sae_out, feature_acts, loss, mse, l1, ghost, aux = tc(x, y, (since_fired > cfg.dead_feature_window).bool())   # :335-344
            with torch.no_grad():                                 # :356-365
                did_fire = (feature_acts > 0).float().sum(-2) > 0
                since_fired += 1
                since_fired[did_fire] = 0
                act_freq += (feature_acts.abs() > 0).float().sum(0)
                l0 = (feature_acts > 0).float().sum(-1).mean()
            loss.backward()
            raw = {n: p.grad.detach().clone() for n, p in tc.named_parameters()}
            gnorm = torch.nn.utils.clip_grad_norm_(tc.parameters(), max_norm=cfg.max_grad_norm)
            tc.remove_gradient_parallel_to_decoder_directions()
            opt.step()
            sched.step()
            rec = dict(lr=lr_now, loss=float(loss.detach()), mse=float(mse.detach()), l1=(None if l1 is None else float(l1.detach())), l0=l0.item(),
                       grad_norm=float(gnorm), sae_out=sae_out.detach().clone())
            if s == 0:
                rec["feature_acts"] = feature_acts.detach().clone()
            if s in (0, 3):
                rec["raw_grads"] = raw
            if s in (0, 2, 4):
                rec["params_after"] = {k_: v.detach().clone() for k_, v in tc.state_dict().items()}
            steps.append(rec)
        path = os.path.join(HERE, f"transcoder_{tag}.pt")
        torch.save(dict(d=d, d_sae=cfg.d_sae, batch=batch, norm=norm, act=act, k=k, l1_coefficient=l1c, skip=skip, lr=cfg.lr,
                        warm_up_steps=cfg.lr_warm_up_steps, total_steps=total_steps, lr_end=cfg.lr / 10, data_seed=13, n_steps=n_steps,
                        init=init, steps=steps, since_fired=since_fired.clone(), act_freq=act_freq.clone()), path)
        print("wrote", path, os.path.getsize(path), "bytes; loss", [round(r["loss"], 4) for r in steps], "l0", [round(r["l0"], 2) for r in steps])


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    import _ref_shims
    _ref_shims.install()
    make()
