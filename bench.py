#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on a B200: HookedViT.run_with_cache images/sec (+ SAE tokens/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload vit|sae]
                    [--dtype fp32|bf16] [--batch B]

Prints ONE JSON line (contract in the task statement):
  value      whole-job throughput with inputs resident in HBM (device-timed, CUDA events, max over ranks)
  e2e        same metric through the public API with HOST (pinned) inputs: H2D of the batch + the call + D2H
             of the model output inside the timed region
  roofline   dominant kernel (MLP-in GEMM with its dual hook-point epilogue), algorithmic flops / CUDA-event time
             against MEASURED_PEAKS.json
  cpu_baseline  the oracle port (oracle/vit_oracle.py) timed on this box's host cores on a bounded sample
--impl reference times the CPU implementation (oracle port; /root/reference does not exist on the GPU box).
A "step" = one run_with_cache over one synthetic batch (cfg #2: CLIP ViT-B/32 geometry, batch 512, all 214 hook points).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vit-prisma_b200"))

import torch  # noqa: E402


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md), through NVML in this process.

    NVML is initialised when the sampler is constructed (before the warm-up steps): starting `nvidia-smi` next to the timed
    loop costs seconds of driver initialisation that stall this process's own launches, and its first sample arrives after a
    short timed region is already over.  The polling thread sleeps between reads; only samples taken between __enter__ and
    __exit__ are reported."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index=0, period_s=0.02):
        self.rows, self.active, self.stop, self.period = [], False, False, period_s
        self.h = self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].strip().isdigit() else index
            self.h, self.nv = pynvml.nvmlDeviceGetHandleByIndex(phys), pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.h = None

    def _pump(self):
        nv = self.nv
        while not self.stop:
            if self.active:
                try:
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), int(mask)))
                except Exception:
                    pass
            time.sleep(self.period)

    def __enter__(self):
        self.active = True
        return self

    def __exit__(self, *exc):
        self.active = False

    def close(self):
        self.stop = True

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({name for _, mask in self.rows for name, bit in self.REASONS if mask & bit})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": getattr(self, "max_mhz", None) if self.h else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml (in-process, polled during the timed region)" if self.h else "unavailable"}


def _dist():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def _cpu_cores():
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota
    (the GPU box shows 128 CPUs but cpu.max grants 16; 128 torch threads on 16 CPUs run 150x slower)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


# ------------------------------------------------------------------ CPU (oracle port) arm
def cpu_vit_images_per_sec(budget_s=12.0, batch=16, threads=None, cfg=None, layer=None):
    """Oracle port on the host cores; ``layer`` = the activation-store call (names_filter one resid_post + stop_at_layer)."""
    from oracle.vit_oracle import CLIP_B32, recipe_state_dict, state_dict_shapes, vit_forward_with_cache
    threads = threads or _cpu_cores()
    torch.set_num_threads(threads)
    cfg = dict(cfg or CLIP_B32)
    kw = {}
    if layer is not None:
        name = f"blocks.{layer}.hook_resid_post"
        kw = dict(names_filter=lambda n: n == name, stop_at_layer=layer + 1)
    sd = recipe_state_dict(state_dict_shapes(cfg), 1234)
    x = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        vit_forward_with_cache(sd, cfg, x, **kw)  # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            vit_forward_with_cache(sd, cfg, x, **kw)
            n += 1
            dt = time.perf_counter() - t0
            if dt >= budget_s or n >= 64:
                break
    what = f"resid_post of layer {layer}, stop_at_layer {layer + 1}" if layer is not None else "all default hook points"
    return {"value": n * batch / dt, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"{n} x run_with_cache(batch {batch}) d_model {cfg['d_model']} x {cfg['n_layers']} layers fp32, {what}, oracle/vit_oracle.py, {dt:.1f}s"}


def run_reference_arm(args):
    """The reference's own CPU path (its restatement, oracle/: /root/reference does not exist on the GPU box) on this box's host
    cores, same metric / unit / workload as the product arm, each step a bounded sample of that workload.  Rank 0 only."""
    world, rank, _ = _dist()
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    threads = _cpu_cores()
    torch.set_num_threads(threads)
    if args.workload == "sae":
        from oracle.sae_oracle import new_adam_state, sae_train_step
        d, F, k = SAE_CFG["d_in"], SAE_CFG["d_in"] * SAE_CFG["expansion"], SAE_CFG["k"]
        p0 = sae_init_params(d, F)
        p = {"W_enc": p0["W_encT"].t().contiguous(), "W_dec": p0["W_dec"], "b_enc": p0["b_enc"], "b_dec": p0["b_dec"]}
        state = new_adam_state(p)
        batch = 1024   # bounded sample of the 4096-token step: the dense products are linear in the token count
        x = sae_pool(batch, d)
        for i in range(max(1, min(warm, 2))):
            sae_train_step(p, state, x, k, 1e-3, i + 1)
        t0 = time.perf_counter()
        for i in range(steps):
            sae_train_step(p, state, x, k, 1e-3, i + 3)
        dt = time.perf_counter() - t0
        v, unit = steps * batch / dt, "tokens/s"
        metric = "SAE training tokens/sec (TopK SAE, d_model=768, dict=768x32, k=32)"
        config = {"workload": "sae_topk_768x24576_k32", "tokens_per_step": batch, "note": "bounded sample of the 4096-token step on host cores"}
        sample = f"{steps} train steps x {batch} tokens, oracle/sae_oracle.py (CPU restatement of the reference train_step)"
    else:
        from oracle.vit_oracle import CLIP_B32, CLIP_L14, recipe_state_dict, state_dict_shapes, vit_forward_with_cache
        l14 = args.model == "l14"
        cfg = dict(CLIP_L14 if l14 else CLIP_B32)
        kw = {}
        if l14:
            name = f"blocks.{args.layer}.hook_resid_post"
            kw = dict(names_filter=lambda n: n == name, stop_at_layer=args.layer + 1)
        sd = recipe_state_dict(state_dict_shapes(cfg), 1234)
        batch = 4 if l14 else 16  # bounded sample of the batch-512 workload: per-image cost is flat in batch on CPU
        x = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(0))
        with torch.no_grad():
            for _ in range(max(1, min(warm, 2))):
                vit_forward_with_cache(sd, cfg, x, **kw)
            t0 = time.perf_counter()
            for _ in range(steps):
                vit_forward_with_cache(sd, cfg, x, **kw)
            dt = time.perf_counter() - t0
        v, unit = steps * batch / dt, "images/s"
        metric = (f"run_with_cache images/sec (CLIP ViT-L/14, blocks.{args.layer}.hook_resid_post, stop_at_layer={args.layer + 1})" if l14
                  else "run_with_cache images/sec (CLIP ViT-B/32, all hook points cached)")
        config = {"workload": f"vit_l14_run_with_cache_resid_post_l{args.layer}" if l14 else "vit_b32_run_with_cache_all_hooks",
                  "batch_per_step": batch, "note": "bounded sample of the product arm's batch on host cores"}
        sample = f"{steps} steps x batch {batch}, oracle/vit_oracle.py (CPU restatement of the reference path)"
    line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": config, "cpu_baseline": {"value": v, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def build_model(dtype, device, cfg=None):
    from oracle.vit_oracle import CLIP_B32, recipe_state_dict
    from vit_prisma.configs.HookedViTConfig import HookedViTConfig
    from vit_prisma.models.base_vit import HookedViT
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = HookedViT(HookedViTConfig(**(cfg or CLIP_B32), dtype=dtype))
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(recipe_state_dict(shapes, 1234))
    return model.to(device, dtype).eval()


def vit_flops_per_image(cfg, stop_at_layer=None):
    """full depth, or L = stop_at_layer without the head term."""
    N = (cfg["image_size"] // cfg["patch_size"]) ** 2
    T = N + 1
    d, H, dh, M, L = cfg["d_model"], cfg["n_heads"], cfg["d_head"], cfg["d_mlp"], cfg["n_layers"]
    CPP = cfg["n_channels"] * cfg["patch_size"] ** 2
    head = 2 * d * cfg["n_classes"]
    if stop_at_layer is not None:
        L, head = stop_at_layer, 0
    return 2 * N * CPP * d + L * (6 * T * d * H * dh + 4 * H * T * T * dh + 2 * T * H * dh * d + 4 * T * d * M) + head


def time_dominant_gemm(model, batch, dtype, iters=10):
    """MLP-in GEMM (+bias, GELU, two hook-point outputs) at the step's shape, alone on the stream, CUDA events."""
    from vit_prisma.b200 import ops
    cfg = model.cfg
    M, K, N = batch * cfg.n_tokens, cfg.d_model, cfg.d_mlp
    mlp = model.blocks[0].mlp
    win, win_lo = mlp.packed_in()
    a = torch.randn(M, K, device="cuda", dtype=dtype)
    a_lo = ops.split_tf32(a) if dtype == torch.float32 else None
    pre = torch.empty(M, N, device="cuda", dtype=dtype)
    post = torch.empty(M, N, device="cuda", dtype=dtype)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    def call():
        ops.gemm(a, win, mlp.b_in, act=cfg.activation_name, a_lo=a_lo, w_lo=win_lo, out0=pre, out1=post)
    for _ in range(3):
        call()
    times = []
    for _ in range(iters):
        flush.zero_()                      # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    es = 4 if dtype == torch.float32 else 2
    return {"ms": ms, "flops": 2.0 * M * N * K, "bytes": float(M * K * es + N * K * es + 2 * M * N * es), "shape": [M, N, K]}


def run_ours(args):
    world, rank, local = _dist()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from oracle.vit_oracle import CLIP_B32, CLIP_L14
    from vit_prisma.b200 import _lib as L
    dtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    l14 = args.model == "l14"
    CFG = CLIP_L14 if l14 else CLIP_B32
    # cfg #4 (ViT part): exactly the call VisionActivationsStore.get_activations makes (activations_store.py:262-270)
    run_kw = dict(names_filter=[f"blocks.{args.layer}.hook_resid_post"], stop_at_layer=args.layer + 1) if l14 else {}
    model = build_model(dtype, dev, CFG)
    B = args.batch
    g = torch.Generator().manual_seed(rank)
    host = torch.randn(B, 3, 224, 224, generator=g).to(dtype).pin_memory()
    x = host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    clocks = ClockSampler(local)
    for _ in range(args.warmup):
        out, cache = model.run_with_cache(x, **run_kw)
        del cache
    barrier()
    n_keys = 0
    launches0 = L.get_lib().pb_launch_count()
    with clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out, cache = model.run_with_cache(x, **run_kw)
            n_keys = len(cache)
            del cache
        e1.record()
        barrier()
        dev_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks.close()
    launches = L.get_lib().pb_launch_count() - launches0
    route = model.last_route

    # ---- end to end: pinned host batch -> H2D -> run_with_cache -> D2H of the model output, every step
    # result read back every step: the model output [B, n_classes]; with stop_at_layer the output is the residual stream
    # (stays on the device for the SAE), so the class-token row of every image [B, d_model] is what crosses PCIe
    out_host = torch.empty((B, CFG["d_model"] if l14 else CFG["n_classes"]), dtype=dtype).pin_memory()
    result = (lambda o: o[:, 0, :]) if l14 else (lambda o: o)
    # (the copy of batch i+1 runs on a side stream under batch i's forward -- vit_prisma.b200.prefetch.DevicePrefetcher, the
    # loader-side helper the package ships; every step's bytes still cross PCIe inside the timed region)
    from vit_prisma.b200.prefetch import DevicePrefetcher
    loader = DevicePrefetcher(None, dev)         # built once, like a DataLoader: the timed region holds per-step work only
    for xd in loader.feed(host for _ in range(max(3, args.warmup))):
        out, cache = model.run_with_cache(xd, **run_kw)
        out_host.copy_(result(out), non_blocking=True)
        del cache
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    marks = []
    for xd in loader.feed(host for _ in range(args.steps)):
        out, cache = model.run_with_cache(xd, **run_kw)
        out_host.copy_(result(out), non_blocking=True)
        del cache
        if os.environ.get("PRISMA_BENCH_DEBUG"):
            marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record()
    e1.record()
    del xd
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    if marks:
        print("e2e per-step ms:", [round(a.elapsed_time(b), 2) for a, b in zip([e0] + marks[:-1], marks)], file=sys.stderr)

    if rank != 0:
        return
    peaks = _peaks()
    imgs = world * B * args.steps
    value = imgs / (dev_ms / 1e3)
    e2e = imgs / (e2e_ms / 1e3)
    kern = time_dominant_gemm(model, B, dtype)
    flops_img = vit_flops_per_image(CFG, run_kw.get("stop_at_layer"))
    cache_b_img = (CFG["d_model"] * ((CFG["image_size"] // CFG["patch_size"]) ** 2 + 1) * 4) if l14 else 38_980_176   # fp32 bytes
    # fp32 mode executes 3 tensor-core passes per algorithmic flop; the roofline counts ALGORITHMIC flops
    achieved = kern["flops"] / (kern["ms"] / 1e3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of this launch from one `ncu --set full` capture (profiles/r01_gemm_fp32_ncu_summary.txt,
    # profiles/r01_gemm_bf16_v3_ncu_summary.txt); algorithmic: A (+ lo plane) + weights read, two M x N outputs written
    traffic = None if l14 else ((180.070400e6 + 580.768512e6) if dtype == torch.float32 else (44.133632e6 + 262.480640e6))
    roof = {"bound": "tensor", "kernel": "k_gemm_tc2 (MLP-in GEMM + bias + GELU, hook_pre/hook_post spill)", "achieved": achieved,
            "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
            "traffic_unit": "B/launch (ncu dram read+write)", "algorithmic_bytes": kern["bytes"],
            "executed_tensor_tflops": achieved * (3 if dtype == torch.float32 else 1),
            "note": ("fp32 = 3 TF32 passes per algorithmic flop at half the bf16 MMA rate: ncu shows the tensor pipe 71.5 % active "
                     "on this launch (profiles/r01_gemm_fp32_ncu_summary.txt); frac is algorithmic flops over the bf16 peak") if dtype == torch.float32 else
                    "ncu: tensor pipe 36.6 % active, issue slots 50.6 % (profiles/r01_gemm_bf16_v3_ncu_summary.txt)",
            "peak_source": peaks["source"] + " bf16 burst (kernel timed alone)", "shape_MNK": kern["shape"], "kernel_ms": kern["ms"],
            "hbm_gbs_of_kernel": kern["bytes"] / (kern["ms"] / 1e3) / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"],
            "passes": 3 if dtype == torch.float32 else 1,
            "step_algorithmic_tflops": flops_img * imgs / (dev_ms / 1e3) / 1e12,
            "step_cache_write_gbs": cache_b_img * (1 if dtype == torch.float32 else 0.5) * imgs / (dev_ms / 1e3) / 1e9}
    cpu = cpu_vit_images_per_sec(batch=4, cfg=CFG, layer=args.layer) if l14 else cpu_vit_images_per_sec()
    es = 4 if dtype == torch.float32 else 2
    metric = (f"run_with_cache images/sec (CLIP ViT-L/14, blocks.{args.layer}.hook_resid_post, stop_at_layer={args.layer + 1})" if l14
              else "run_with_cache images/sec (CLIP ViT-B/32, all hook points cached)")
    line = {"metric": metric, "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"vit_l14_run_with_cache_resid_post_l{args.layer}" if l14 else "vit_b32_run_with_cache_all_hooks",
                       "model": ("CLIP ViT-L/14" if l14 else "CLIP ViT-B/32") + " geometry, seeded synthetic weights",
                       "batch_per_gpu": B, "global_batch": B * world, "hook_points_cached": n_keys, "route": route,
                       "gemm": "tcgen05 3xTF32" if dtype == torch.float32 else "tcgen05 bf16",
                       "cache_bytes_per_image": int(cache_b_img * es / 4), "l2": "working set (activations of one layer >> 126 MB) larger than L2",
                       "parallelism": f"dp{world} (images sharded, no collective)"},
            "clocks": clocks.summary(), "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(host.numel() * host.element_size()),
                    "d2h_bytes_per_step": int(out_host.numel() * out_host.element_size()), "ms_per_step": e2e_ms / args.steps,
                    "overlap": "H2D of step i+1 on a copy stream under step i (DevicePrefetcher, depth 2)"},
            "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ SAE workload (cfg #3: d=768, F=768*32, k=32, 4096 tokens/step, fp32)
SAE_CFG = dict(d_in=768, expansion=32, k=32, batch=4096)


def sae_init_params(d, F, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    W_dec = torch.randn(F, d, generator=g)
    W_dec /= W_dec.norm(dim=1, keepdim=True)
    W_encT = torch.randn(F, d, generator=g)
    W_encT /= W_encT.norm(dim=0, keepdim=True) + 1e-12      # reference: rows of W_enc [d,F] unit-norm
    return dict(W_encT=W_encT.to(device), W_dec=W_dec.to(device), b_enc=torch.zeros(F, device=device), b_dec=torch.zeros(d, device=device))


def sae_pool(tokens, d, seed=0):
    g = torch.Generator().manual_seed(seed)
    off = torch.randn(d, generator=g)
    return torch.randn(tokens, d, generator=g) * 2.0 + off


def cpu_sae_tokens_per_sec(budget_s=12.0, threads=None, batch=1024):
    """Oracle port of the reference train_step (dense autograd-equivalent formulas) on host cores, bounded sample."""
    from oracle.sae_oracle import new_adam_state, sae_train_step
    threads = threads or _cpu_cores()
    torch.set_num_threads(threads)
    d, F, k = SAE_CFG["d_in"], SAE_CFG["d_in"] * SAE_CFG["expansion"], SAE_CFG["k"]
    p0 = sae_init_params(d, F)
    p = {"W_enc": p0["W_encT"].t().contiguous(), "W_dec": p0["W_dec"], "b_enc": p0["b_enc"], "b_dec": p0["b_dec"]}
    state = new_adam_state(p)
    x = sae_pool(batch, d)
    sae_train_step(p, state, x, k, 1e-3, 1)
    n, t0 = 0, time.perf_counter()
    while True:
        sae_train_step(p, state, x, k, 1e-3, n + 2)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 16:
            break
    return {"value": n * batch / dt, "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"{n} train steps x {batch} tokens (of the 4096-token step), d=768 F=24576 k=32 fp32, oracle/sae_oracle.py, {dt:.1f}s"}


def run_sae(args):
    world, rank, local = _dist()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from vit_prisma.b200 import _lib as L
    from vit_prisma.b200.sae_engine import SaeStepEngine, unit_norm_rows_
    d, F, k, Bt = SAE_CFG["d_in"], SAE_CFG["d_in"] * SAE_CFG["expansion"], SAE_CFG["k"], SAE_CFG["batch"]
    p = sae_init_params(d, F, device=dev)
    if world > 1:      # data parallel over NVLink peer memory: Bt tokens per GPU, one global step (csrc/p2p.cu)
        from vit_prisma.b200.p2p import P2PGroup, SaeDPEngine
        eng = SaeDPEngine(P2PGroup(rank, world, dev), p["W_encT"], p["W_dec"], p["b_enc"], p["b_dec"], k=k)
    else:
        eng = SaeStepEngine(p["W_encT"], p["W_dec"], p["b_enc"], p["b_dec"], k=k)
    unit_norm_rows_(eng.W_dec)
    eng.refresh_lo()
    pool_host = sae_pool(Bt * 16, d, seed=rank).pin_memory()
    pool = pool_host.to(dev)
    eng.b_dec.copy_(sae_pool(Bt * 16, d, seed=0).mean(0).to(dev))     # identical b_dec init on every rank
    since_fired, act_freq = torch.zeros(F, device=dev), torch.zeros(F, device=dev)
    lr = 1e-3

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def batch(i):
        j = (i % 16) * Bt
        return pool[j:j + Bt]

    clocks = ClockSampler(local)
    for i in range(args.warmup):
        eng.train_step(batch(i), lr, since_fired, act_freq)
    barrier()
    l0 = L.get_lib().pb_launch_count()
    with clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for i in range(args.steps):
            eng.train_step(batch(i), lr, since_fired, act_freq)
        e1.record()
        host_enqueue_ms = 1e3 * (time.perf_counter() - t_host) / args.steps
        barrier()
        dev_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks.close()
    launches = L.get_lib().pb_launch_count() - l0
    # e2e: pinned host tokens -> H2D -> step -> D2H of the step scalars (mse, l0, grad norm, ...)
    sc_host = torch.empty(8).pin_memory()
    from vit_prisma.b200.prefetch import DevicePrefetcher
    host_batches = lambda n: (pool_host[(i % 16) * Bt:(i % 16 + 1) * Bt] for i in range(n))
    loader = DevicePrefetcher(None, dev)
    for xin in loader.feed(host_batches(max(3, args.warmup))):
        sc_host.copy_(eng.train_step(xin, lr, since_fired, act_freq), non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    for xin in loader.feed(host_batches(args.steps)):
        sc_host.copy_(eng.train_step(xin, lr, since_fired, act_freq), non_blocking=True)
    e1.record()
    e2e_host_ms = 1e3 * (time.perf_counter() - t_host) / args.steps
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    if rank != 0:
        return
    # per-stage device times (one extra instrumented step) -> dominant kernel for the roofline
    import ctypes as C
    lib, st = L.get_lib(), torch.cuda.current_stream().cuda_stream
    stages = {}

    def timed(name, fn, reps=5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        stages[name] = a.elapsed_time(b) / reps

    x = batch(0).contiguous()
    timed("encode_topk (prep + encoder GEMM 3xTF32 + topk)", lambda: eng.encode_topk(x))
    eng.scalars.zero_(); eng.step_count += 1
    s = eng._desc(x, training=True, lr=lr, since_fired=since_fired, act_freq=act_freq, want_out=False)
    timed("decode (sparse decode + loss + d_hidden)", lambda: L.check(lib.pb_sae_decode(C.byref(s), st)))
    timed("backward (csc + per-feature grads + norm)", lambda: (eng.scalars.zero_(), L.check(lib.pb_sae_backward(C.byref(s), st))))
    timed("adam (clip + projection + Adam + renorm)", lambda: L.check(lib.pb_sae_adam(C.byref(s), st)))
    peaks = _peaks()
    tokens = world * Bt * args.steps
    value = tokens / (dev_ms / 1e3)
    step_bytes = eng.algorithmic_bytes(Bt)
    timed("encoder GEMM alone (hidden_pre = sae_in @ W_enc + b_enc, 3xTF32)", lambda: eng._encoder_gemm(Bt))
    adam_bytes = 60 * d * F            # per matrix element pair: g,p,m,v read (16 B) + p,m,v write (12 B) (x2) + tf32 residual write (4 B)
    adam_gbs = adam_bytes / (stages["adam (clip + projection + Adam + renorm)"] / 1e3) / 1e9
    gemm_ms = stages["encoder GEMM alone (hidden_pre = sae_in @ W_enc + b_enc, 3xTF32)"]
    gemm_tflops = 2.0 * Bt * d * F / (gemm_ms / 1e3) / 1e12
    # dominant kernel of the step (41 % of the launch time, profiles/r01_launches_sae_v2.txt): the fp32-grade encoder GEMM.
    # The step as a whole is the HBM-bound object SURVEY 8d describes; its bytes and the Adam kernel's are reported beside it.
    roof = {"bound": "tensor", "kernel": "k_gemm_tc2<float,3,128,3,4> (encoder GEMM, 3 TF32 passes, m-fastest raster)", "achieved": gemm_tflops,
            "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": gemm_tflops / peaks["bf16_tflops"], "traffic": None,
            "peak_source": peaks["source"] + " bf16 burst (kernel timed alone)", "kernel_ms": gemm_ms, "passes": 3,
            "executed_tensor_tflops": 3 * gemm_tflops, "shape_MNK": [Bt, F, d],
            "hbm_kernel": {"kernel": "k_sae_adam_rows (clip + decoder-parallel-grad removal + Adam + row renorm)", "achieved": adam_gbs,
                           "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": adam_gbs / peaks["hbm_gbs"],
                           "algorithmic_bytes_per_launch": adam_bytes, "traffic": 604.774912e6 + 474.265088e6,
                           "traffic_source": "ncu dram read+write, profiles/r01_sae_step_ncu_summary.txt"},
            "step_algorithmic_bytes": step_bytes, "step_hbm_gbs": step_bytes / (dev_ms / args.steps / 1e3) / 1e9,
            "step_hbm_frac": step_bytes / (dev_ms / args.steps / 1e3) / 1e9 / peaks["hbm_gbs"], "stage_ms": stages}
    cpu = cpu_sae_tokens_per_sec()
    line = {"metric": "SAE training tokens/sec (TopK SAE, d_model=768, dict=768x32, k=32)", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "sae_topk_train_step", "d_in": d, "d_sae": F, "k": k, "tokens_per_step_per_gpu": Bt,
                       "encoder_gemm": "tcgen05 3xTF32", "normalize_activations": "layer_norm", "max_grad_norm": 1.0,
                       "l2": "working set (2 x 75 MB weights + 2 x 150 MB Adam state + 150 MB grads + 403 MB hidden_pre) larger than L2",
                       "parallelism": f"dp{world}" + (" (NVLink peer-memory reduce-scatter + sharded Adam + all-gather, no NCCL on the data path; "
                                                       f"{int((world - 1) / world * (2 * d * F * 4 * 2 + d * F * 4) / 1e6)} MB over NVLink per GPU per step)" if world > 1 else "")},
            "clocks": clocks.summary(), "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms,
            "e2e": {"value": tokens / (e2e_ms / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": Bt * d * 4, "d2h_bytes_per_step": 32,
                    "ms_per_step": e2e_ms / args.steps, "host_enqueue_ms_per_step": e2e_host_ms,
                    "overlap": "H2D of step i+1 on a copy stream under step i (DevicePrefetcher, depth 2)"},
            "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vit", choices=["vit", "sae"])
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--model", default="b32", choices=["b32", "l14"], help="l14 = cfg #4: ViT-L/14 with the activation store's names_filter / stop_at_layer")
    ap.add_argument("--layer", type=int, default=22, help="hook_resid_post layer cached by --model l14")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    try:
        if args.workload == "sae":
            return run_sae(args)
        run_ours(args)
    finally:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
