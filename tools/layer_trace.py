"""Layer-by-layer parity trace: product (GPU) vs bf16-emulating oracle.  Debug tool (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dm
from oracle import init as OI
from oracle.ops import Emu
from oracle.unet import UNet as OUNet

S, C_lat, B = 8, 8, 2
net = b200dm.build_model(S, C_lat, [64, 128, 256], [False, False, True, True])
ou = OUNet(S, C_lat, [64, 128, 256], [False, False, True, True], first_conv_channels=64)
P = OI.make_params(ou.spec(), 0, "stress")
net.set_weights(P)
net.compile(B, 50)
x = OI.normal((B, S, S, S, C_lat), 1)
tt = torch.full((B,), 37)
y = net([x.cuda(), tt])
tr = {}
ref = ou.forward(P, x, tt, emu=Emu(True, tr))
tr32 = {}
ref32 = ou.forward(P, x, tt, emu=Emu(False, tr32))
for k, v in net.prog.outputs.items():
    if k in tr:
        a = v.float().cpu().reshape(tr[k].shape)
        e = ((a - tr[k]).norm() / tr[k].norm()).item()
        e32 = ((a - tr32[k]).norm() / tr32[k].norm()).item()
        ee = ((tr[k] - tr32[k]).norm() / tr32[k].norm()).item()
        print(f"{k:28s} prod-emu {e:.3e}  prod-fp32 {e32:.3e}  emu-fp32 {ee:.3e}")
print("final", ((y.cpu() - ref).norm() / ref.norm()).item(), ((y.cpu() - ref32).norm() / ref32.norm()).item())
