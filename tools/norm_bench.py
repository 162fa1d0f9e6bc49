"""Time the fused BN(+swish)(+concat) pass (CUDA events, rotating buffers larger than L2).  ncu target for the HBM-bound kernel.
   python tools/norm_bench.py B S C0 C1 [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dm
from b200dm import ops

a = [int(v) for v in sys.argv[1:]]
B, S, c0, c1 = a[:4]
iters = a[4] if len(a) > 4 else 20
dev = torch.device("cuda", 0)
NBUF = 4
C = c0 + c1
sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
bufs = []
for i in range(NBUF):
    x0 = torch.randn(B, S, S, S, c0, device=dev).bfloat16()
    x1 = torch.randn(B, S, S, S, c1, device=dev).bfloat16() if c1 else None
    bufs.append((x0, x1, torch.empty(B, S, S, S, C, device=dev, dtype=torch.bfloat16)))
for x0, x1, y in bufs:
    ops.norm_act(x0, sc, sh, act="silu", x1=x1, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    x0, x1, y = bufs[i % NBUF]
    ops.norm_act(x0, sc, sh, act="silu", x1=x1, out=y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
nbytes = 2.0 * B * S ** 3 * C * 2
print(f"norm_act B{B} {S}^3 {c0}+{c1}: {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s (algorithmic read+write)")
