"""GPU probe: does an H2D copy on a side stream overlap run_with_cache? (tuning aid; not a bench value)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit-prisma_b200"))
import torch
import bench

dtype = torch.float32 if (len(sys.argv) < 2 or sys.argv[1] == "fp32") else torch.bfloat16
dev = torch.device("cuda", 0)
model = bench.build_model(dtype, dev)
B = 512
host = torch.randn(B, 3, 224, 224).to(dtype).pin_memory()
bufs = [torch.empty_like(host, device=dev) for _ in range(2)]
cs = torch.cuda.Stream()
def ev(): return torch.cuda.Event(enable_timing=True)

def fwd(x):
    out, cache = model.run_with_cache(x); del cache; return out

for _ in range(3): fwd(bufs[0])
torch.cuda.synchronize()
# (a) copy alone
a, b = ev(), ev(); a.record(); bufs[0].copy_(host, non_blocking=True); b.record(); torch.cuda.synchronize()
print(f"copy alone {a.elapsed_time(b):.2f} ms  ({host.numel()*host.element_size()/a.elapsed_time(b)/1e6:.1f} GB/s)")
# (b) compute alone
a, b = ev(), ev(); a.record(); fwd(bufs[0]); b.record(); torch.cuda.synchronize()
print(f"forward alone {a.elapsed_time(b):.2f} ms")
# (c) forward with a concurrent copy on the side stream
for rep in range(3):
    a, b, c0, c1 = ev(), ev(), ev(), ev()
    torch.cuda.synchronize()
    a.record()
    with torch.cuda.stream(cs):
        c0.record(cs); bufs[1].copy_(host, non_blocking=True); c1.record(cs)
    fwd(bufs[0]); b.record(); torch.cuda.synchronize()
    print(f"overlapped: forward {a.elapsed_time(b):.2f} ms, copy {c0.elapsed_time(c1):.2f} ms")
# (d) host-side enqueue time of one forward (is the CPU the limiter?)
torch.cuda.synchronize(); t0 = time.perf_counter(); fwd(bufs[0]); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0):.2f} ms, until done {1e3*(t2-t0):.2f} ms")
# (e) the two e2e loops of bench.py, 5 steps each, interleaved twice
from vit_prisma.b200.prefetch import DevicePrefetcher
out_host = torch.empty((B, 512), dtype=dtype).pin_memory()
def serial(n):
    for _ in range(n):
        xd = host.to(dev, non_blocking=True); out = fwd(xd); out_host.copy_(out, non_blocking=True)
def piped(n):
    for xd in DevicePrefetcher((host for _ in range(n)), dev):
        out = fwd(xd); out_host.copy_(out, non_blocking=True)
for name, fn in (("serial", serial), ("prefetch", piped), ("serial", serial), ("prefetch", piped)):
    fn(2); torch.cuda.synchronize()
    a, b = ev(), ev(); t0 = time.perf_counter(); a.record(); fn(5); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b)/5:.2f} ms/step (host loop {1e3*(t1-t0)/5:.2f} ms/step)")
