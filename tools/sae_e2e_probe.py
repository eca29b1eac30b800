"""GPU probe: SAE e2e loop variants (tuning aid; not a bench value)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit-prisma_b200"))
import torch
import bench
from vit_prisma.b200.sae_engine import SaeStepEngine, unit_norm_rows_
from vit_prisma.b200.prefetch import DevicePrefetcher
dev = torch.device("cuda", 0)
d, F, k, Bt = 768, 768 * 32, 32, 4096
p = bench.sae_init_params(d, F, seed=0, device=dev)
eng = SaeStepEngine(p["W_encT"], p["W_dec"], p["b_enc"], p["b_dec"], k=k)
unit_norm_rows_(eng.W_dec); eng.refresh_lo()
pool_host = bench.sae_pool(Bt * 16, d, seed=0).pin_memory(); pool = pool_host.to(dev)
sf, af = torch.zeros(F, device=dev), torch.zeros(F, device=dev)
sc_host = torch.empty(8).pin_memory()
def ev(): return torch.cuda.Event(enable_timing=True)
hb = lambda n: (pool_host[(i % 16) * Bt:(i % 16 + 1) * Bt] for i in range(n))
xin = torch.empty(Bt, d, device=dev)
def resident(n):
    for i in range(n): eng.train_step(pool[(i % 16) * Bt:(i % 16 + 1) * Bt], 1e-3, sf, af)
def resident_d2h(n):
    for i in range(n): sc_host.copy_(eng.train_step(pool[(i % 16) * Bt:(i % 16 + 1) * Bt], 1e-3, sf, af), non_blocking=True)
def serial(n):
    for h in hb(n):
        xin.copy_(h, non_blocking=True); sc_host.copy_(eng.train_step(xin, 1e-3, sf, af), non_blocking=True)
def serial_nod2h(n):
    for h in hb(n):
        xin.copy_(h, non_blocking=True); eng.train_step(xin, 1e-3, sf, af)
def piped(n):
    for x in DevicePrefetcher(hb(n), dev): sc_host.copy_(eng.train_step(x, 1e-3, sf, af), non_blocking=True)
def piped_nod2h(n):
    for x in DevicePrefetcher(hb(n), dev): eng.train_step(x, 1e-3, sf, af)
for name, fn in (("resident", resident), ("resident+d2h", resident_d2h), ("serial", serial), ("serial no d2h", serial_nod2h),
                 ("prefetch", piped), ("prefetch no d2h", piped_nod2h), ("resident", resident)):
    fn(3); torch.cuda.synchronize()
    a, b = ev(), ev(); t0 = time.perf_counter(); a.record(); fn(20); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name:16s}: {a.elapsed_time(b)/20:.3f} ms/step (host {1e3*(t1-t0)/20:.3f})", flush=True)
