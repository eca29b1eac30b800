"""GPU probe: fused hooked attention at the ViT-B/32 batch-512 shape (for ncu / event timing; not a bench value)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit-prisma_b200"))
import torch
from vit_prisma.b200 import ops

dtype = torch.float32 if (len(sys.argv) > 1 and sys.argv[1] == "fp32") else torch.bfloat16
B, T, H, dh = 512, 50, 12, 64
q, k, v = (torch.randn(B, T, H, dh, device="cuda").to(dtype) for _ in range(3))
for _ in range(3): ops.attention(q, k, v, 8.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attention(q, k, v, 8.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
es = 2 if dtype == torch.bfloat16 else 4
byts = (4 * B * T * H * dh + 2 * B * H * T * T) * es
print(f"{dtype} attention B={B} T={T} H={H}: {ms*1e3:.1f} us/launch (incl. 3 output allocations), {byts/ms/1e6:.0f} GB/s algorithmic")
