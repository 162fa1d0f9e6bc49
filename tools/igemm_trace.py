"""clock64 timeline of CTA 0 of the per-tap GEMM kernel for one small launch (tuning aid).  args: B D H W C0 COUT KSIZE"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import b200dm
from b200dm import ops, _lib
B, D, H, W, c0, cout, k = [int(v) for v in sys.argv[1:8]]
dev = torch.device("cuda", 0)
w = torch.randn(k, k, k, c0, cout) * 0.05
desc = ops.make_conv_desc(0, B, (D, H, W), c0, 0, cout, k, 1, use_halo=-1)
wp = ops.pack_conv_weights(desc, w).to(dev)
x0 = torch.randn(B, D, H, W, c0, device=dev).bfloat16()
y = torch.empty(B, D, H, W, cout, device=dev, dtype=torch.bfloat16)
plan = ops.ConvPlan(desc, x0, wp, y, bias=torch.zeros(cout, device=dev))
for _ in range(3):
    plan.run()
tr = torch.zeros(4 * 2048, dtype=torch.int64, device=dev)
_lib.check(_lib.lib().b200dm_conv_plan_set_trace(plan.h, C.c_void_p(tr.data_ptr())))
torch.cuda.synchronize()
plan.run()
torch.cuda.synchronize()
t = tr.cpu().view(4, 2048)
names = {40: "start", 41: "setup done", 42: "pdl_wait done", 43: "producer: first stage free", 44: "mma: stage full", 45: "epi: wait acc",
         46: "epi: acc ready", 47: "epi: done", 48: "exit"}
ev = []
for region in range(4):
    for v in t[region]:
        v = int(v)
        if v == 0:
            break
        ev.append((v >> 8, names.get(v & 0xff, str(v & 0xff))))
ev.sort()
t0 = ev[0][0]
for c, n in ev:
    print(f"{c - t0:8d}  {n}")
print(plan.info)
