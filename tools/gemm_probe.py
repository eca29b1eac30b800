"""GPU probe: time GEMM variants with CUDA events (for kernel tuning; not a bench value)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit-prisma_b200"))
import torch
from vit_prisma.b200 import ops, _lib as L

def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def probe(M, N, K, dtype, variants):
    a = torch.randn(M, K, device="cuda", dtype=dtype); w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dtype)
    b = torch.randn(N, device="cuda", dtype=dtype); r = torch.randn(M, N, device="cuda", dtype=dtype)
    o0 = torch.empty(M, N, device="cuda", dtype=dtype); o1 = torch.empty(M, N, device="cuda", dtype=dtype)
    lo = dict(a_lo=ops.split_tf32(a), w_lo=ops.split_tf32(w)) if dtype == torch.float32 else {}
    for name in variants:
        if name == "pre": fn = lambda: ops.gemm(a, w, b, out0=o0, impl=L.GEMM_TC, **lo)
        elif name == "pre_nobias": fn = lambda: ops.gemm(a, w, None, out0=o0, impl=L.GEMM_TC, **lo)
        elif name == "pre+gelu": fn = lambda: ops.gemm(a, w, b, act="gelu", out0=o0, out1=o1, impl=L.GEMM_TC, **lo)
        elif name == "gelu_only": fn = lambda: ops.gemm(a, w, b, act="gelu", want_pre=False, out1=o1, impl=L.GEMM_TC, **lo)
        elif name == "pre+resid": fn = lambda: ops.gemm(a, w, b, residual=r, out0=o0, out1=o1, impl=L.GEMM_TC, **lo)
        elif name == "simt": fn = lambda: ops.gemm(a, w, b, out0=o0, impl=L.GEMM_SIMT)
        elif name == "torch": fn = lambda: torch.matmul(a, w.t(), out=o0)
        ms = timeit(fn)
        err = ""
        if name in ("pre", "pre+resid") and os.environ.get("PROBE_CHECK"):       # numerics of the variant under test vs fp32 torch
            ref = a.float() @ w.float().t() + b.float()
            got = o0.float()
            err = f"  max rel err {((got - ref).abs().max() / ref.abs().max()).item():.2e}"
        print(f"{str(dtype):15s} M={M} N={N} K={K} {name:12s} {ms:8.3f} ms  {2.0*M*N*K/ms/1e9:8.1f} TFLOP/s{err}", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "one":   # single launch for ncu
        probe(25600, 3072, 768, torch.bfloat16, ["pre+gelu"]); sys.exit(0)
    if which == "one32":
        probe(25600, 3072, 768, torch.float32, ["pre+gelu"]); sys.exit(0)
    probe(25600, 3072, 768, torch.bfloat16, ["pre_nobias", "pre", "pre+gelu", "gelu_only", "torch"])
    probe(25600, 768, 3072, torch.bfloat16, ["pre", "pre+resid", "torch"])
    probe(25600, 2304, 768, torch.bfloat16, ["pre", "torch"])
    probe(25600, 3072, 768, torch.float32, ["pre", "pre+gelu", "torch"])
    probe(25600, 768, 768, torch.bfloat16, ["pre", "pre+resid", "torch"])
    probe(25600, 768, 3072, torch.float32, ["pre", "pre+resid"])
    probe(25600, 2304, 768, torch.float32, ["pre"])
    probe(25600, 768, 768, torch.float32, ["pre", "pre+resid"])
    probe(4096, 24576, 768, torch.float32, ["pre"])
    if which != "cpu":
        sys.exit(0)
    # host-thread calibration for the CPU baseline
    from oracle.vit_oracle import CLIP_B32, recipe_state_dict, state_dict_shapes, vit_forward_with_cache
    sd = recipe_state_dict(state_dict_shapes(CLIP_B32), 1234); x = torch.randn(8, 3, 224, 224)
    for th in (4, 8, 16, 32, 64):
        torch.set_num_threads(th)
        with torch.no_grad():
            vit_forward_with_cache(sd, dict(CLIP_B32), x)
            t0 = time.perf_counter(); vit_forward_with_cache(sd, dict(CLIP_B32), x); dt = time.perf_counter() - t0
        print(f"cpu oracle threads={th}: {8/dt:.1f} img/s", flush=True)
    print("sched_getaffinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
    try: print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
    except Exception as e: print("cpu.max n/a", e)
