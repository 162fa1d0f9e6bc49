"""Time one conv configuration (CUDA events, L2-cold by rotating buffers).  Used for tuning and as the ncu target.
   python tools/conv_bench.py B D H W C0 C1 COUT [mode] [use_halo] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dm
from b200dm import ops, _lib

a = [int(v) for v in sys.argv[1:]]
B, D, H, W, c0, c1, cout = a[:7]
mode = a[7] if len(a) > 7 else 0
use_halo = a[8] if len(a) > 8 else 0
iters = a[9] if len(a) > 9 else 20
ksize = a[10] if len(a) > 10 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
NBUF = 4
w = torch.randn(ksize, ksize, ksize, c0 + c1, cout) * 0.05
desc = ops.make_conv_desc(mode, B, (D, H, W), c0, c1, cout, ksize, 1, use_halo=use_halo)
wp = ops.pack_conv_weights(desc, w).to(dev)
bias = torch.zeros(cout, device=dev)
plans = []
od, oh, ow = ops.conv_out_shape(mode, (D, H, W), 1)
for i in range(NBUF):
    x0 = torch.randn(B, D, H, W, c0, device=dev).bfloat16()
    x1 = torch.randn(B, D, H, W, c1, device=dev).bfloat16() if c1 else None
    y = torch.empty(B, od, oh, ow, cout, device=dev, dtype=torch.bfloat16)
    plans.append(ops.ConvPlan(desc, x0, wp, y, x1=x1, bias=bias))
for p in plans:
    p.run()
torch.cuda.synchronize()
assert _lib.debug_flag() == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    plans[i % NBUF].run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = plans[0].flops
print(f"conv B{B} {D}x{H}x{W} {c0}+{c1}->{cout} k{ksize} mode{mode} halo{use_halo}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (algorithmic)")
