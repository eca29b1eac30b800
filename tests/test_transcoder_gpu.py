"""Transcoder (reference sae/transcoder.py:6-116) on the GPU against fixtures made by the UNMODIFIED reference class
(tests/golden/make_golden_transcoder.py): five optimizer steps each for ReLU + L1 + skip connection ("t") and TopK without skip ("u")."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.sae_oracle import lr_multiplier  # noqa: E402
from tests.util import assert_close, load_golden  # noqa: E402


def _pair(gold):
    g = torch.Generator().manual_seed(gold["data_seed"])
    n, d = gold["batch"] * gold["n_steps"], gold["d"]
    x = torch.randn(n, d, generator=g) * 2.0 + torch.randn(d, generator=g)
    M = torch.randn(d, d, generator=g) / d ** 0.5
    y = torch.tanh(x @ M) * 1.5 + 0.3 * torch.randn(n, d, generator=g) + torch.randn(d, generator=g)
    return x, y


def _module(gold, tag):
    from vit_prisma.sae.config import VisionModelSAERunnerConfig
    from vit_prisma.sae.transcoder import Transcoder
    cfg = VisionModelSAERunnerConfig(d_in=gold["d"], d_out=gold["d"], expansion_factor=gold["d_sae"] // gold["d"], activation_fn_str=gold["act"],
                                     activation_fn_kwargs=({"k": gold["k"]} if gold["act"] == "topk" else {}), l1_coefficient=gold["l1_coefficient"],
                                     is_transcoder=True, transcoder_with_skip_connection=gold["skip"], _device="cuda", _dtype="float32",
                                     normalize_activations=gold["norm"], b_dec_init_method="zeros", lr=gold["lr"], lr_warm_up_steps=gold["warm_up_steps"],
                                     train_batch_size=gold["batch"], max_grad_norm=1.0, log_to_wandb=False, n_checkpoints=0,
                                     checkpoint_path="/tmp/prisma_b200_unused", use_ghost_grads=False, verbose=False)
    tc = Transcoder(cfg)
    assert list(tc.state_dict()) == list(gold["init"]), "parameter names / order must match the reference module"
    tc.load_state_dict(gold["init"])
    return cfg, tc


@pytest.mark.parametrize("tag", ["t", "u"])
def test_transcoder_training_matches_reference_fixture(tag):
    from vit_prisma.sae.train_sae import FusedAdamHandle, FusedSchedule, VisionSAETrainer
    from vit_prisma.sae.training.get_scheduler import lr_multiplier_fn
    gold = load_golden(f"transcoder_{tag}.pt")
    x_all, y_all = _pair(gold)
    cfg, _ = _module(gold, tag)
    trainer = VisionSAETrainer(cfg, model=None, dataset=None, activations_store=object())
    tc = trainer.sparse_coder
    tc.load_state_dict(gold["init"])
    F, B = gold["d_sae"], gold["batch"]
    since, freq = torch.zeros(F, device="cuda"), torch.zeros(F, device="cuda")
    opt = FusedAdamHandle(gold["lr"])
    sched = FusedSchedule(opt, gold["lr"], lr_multiplier_fn("cosineannealingwarmup", warm_up_steps=gold["warm_up_steps"],
                                                            training_steps=gold["total_steps"], lr_end=gold["lr_end"]))
    n_frac = 0
    for s, rec in enumerate(gold["steps"]):
        assert abs(opt.param_groups[0]["lr"] - rec["lr"]) <= 1e-12 + 1e-9 * rec["lr"]
        assert abs(rec["lr"] - gold["lr"] * lr_multiplier(s, gold["warm_up_steps"], gold["total_steps"], gold["lr_end"])) < 1e-12
        pair = torch.stack([x_all[s * B:(s + 1) * B], y_all[s * B:(s + 1) * B]], dim=1).cuda()          # [B, 2, d]: train_sae.py:299-301
        loss, mse, l1, l0, freq, since, n_frac = trainer.train_step(tc, opt, sched, freq, since, n_frac, pair, s, s * B)
        eng = tc.step_engine()
        terms = eng.loss_terms(B)
        assert abs(terms["mse"] - rec["mse"]) <= 1e-4 * abs(rec["mse"]), (s, terms["mse"], rec["mse"])
        assert abs(float(loss) - rec["loss"]) <= 1e-4 * abs(rec["loss"]), (s, float(loss), rec["loss"])
        assert abs(terms["grad_norm"] - rec["grad_norm"]) <= 2e-4 * rec["grad_norm"], (s, terms["grad_norm"], rec["grad_norm"])
        assert abs(float(l0) - rec["l0"]) <= 1e-3 * max(rec["l0"], 1.0)
        if rec["l1"] is None:
            assert l1 is None
        else:
            assert abs(float(l1) - rec["l1"]) <= 1e-4 * abs(rec["l1"])
        if "raw_grads" in rec:
            g = rec["raw_grads"]
            assert_close(eng.gW_dec.cpu(), g["W_dec"], 2e-4, "dL/dW_dec")
            assert_close(eng.gW_encT.t().cpu(), g["W_enc"], 2e-4, "dL/dW_enc")
            assert_close(eng.gb_enc.cpu(), g["b_enc"], 2e-4, "dL/db_enc")
            assert_close(eng.gb_dec.cpu(), g["b_dec"], 2e-4, "dL/db_dec")
            assert_close(eng.gb_dec_out.cpu(), g["b_dec_out"], 2e-4, "dL/db_dec_out")
            if gold["skip"]:
                assert_close(eng.gW_skip.cpu(), g["W_skip"], 2e-4, "dL/dW_skip")
        if "params_after" in rec:
            ref = rec["params_after"]
            sd = tc.state_dict()
            ref_dec = ref["W_dec"] / ref["W_dec"].norm(dim=1, keepdim=True)     # the reference renormalises at its next step
            assert_close(sd["W_dec"].cpu(), ref_dec, 1e-4, f"step {s} W_dec")
            for name in ("W_enc", "b_dec", "b_dec_out") + (("W_skip",) if gold["skip"] else ()):
                assert_close(sd[name].cpu(), ref[name], 1e-4, f"step {s} {name}")
            assert_close(sd["b_enc"].cpu(), ref["b_enc"], 1e-4 if ref["b_enc"].abs().max() > 1e-3 else 1e-2, f"step {s} b_enc")
    assert torch.equal(since.cpu(), gold["since_fired"]) and torch.equal(freq.cpu(), gold["act_freq"])


@pytest.mark.parametrize("tag", ["t", "u"])
def test_transcoder_forward_tuple_matches_reference_fixture(tag):
    gold = load_golden(f"transcoder_{tag}.pt")
    x_all, y_all = _pair(gold)
    _, tc = _module(gold, tag)
    tc.set_decoder_norm_to_unit_norm()
    B, rec = gold["batch"], gold["steps"][0]
    seen = []
    tc.add_hook("hook_hidden_post", lambda t, hook: seen.append(tuple(t.shape)))
    out = tc(x_all[:B].cuda(), y_all[:B].cuda())
    assert len(out) == 7 and seen == [(B, gold["d_sae"])]
    assert_close(out[0].cpu(), rec["sae_out"], 1e-4, "sae_out")
    assert_close(out[1].cpu(), rec["feature_acts"], 1e-4, "feature_acts")
    assert abs(float(out[3]) - rec["mse"]) <= 1e-4 * rec["mse"] and abs(float(out[2]) - rec["loss"]) <= 1e-4 * rec["loss"]
    assert (out[4] is None) == (rec["l1"] is None)
    sae_in, feats = tc.encode(x_all[:B].cuda())
    assert_close(feats.cpu(), rec["feature_acts"], 1e-4, "encode feature_acts")
    assert tc.decode(feats).shape == (B, gold["d"])
