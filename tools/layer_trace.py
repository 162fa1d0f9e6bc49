"""Layer-by-layer parity trace: product (GPU) vs the 16-bit-emulating oracle and the fp32 oracle.  Debug tool (GPU box).

    python tools/layer_trace.py [S] [C_lat] [B] [has_attention, e.g. 1,0,1] [cond 0|1]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dm
from b200dm import _lib
from oracle import init as OI
from oracle.ops import Emu
from oracle.unet import UNet as OUNet

a = sys.argv[1:]
S, C_lat, B = int(a[0]) if a else 8, int(a[1]) if len(a) > 1 else 8, int(a[2]) if len(a) > 2 else 2
has = [bool(int(v)) for v in a[3].split(",")] if len(a) > 3 else [False, False, True, True]
cond = bool(int(a[4])) if len(a) > 4 else False
net = b200dm.build_model(S, C_lat, [64, 128, 256], has, context_dim=1 if cond else None)
ou = OUNet(S, C_lat, [64, 128, 256], has, first_conv_channels=32 if cond else 64, conditional=cond)
P = OI.make_params(ou.spec(), 0, "stress")
net.set_weights(P)
net.compile(B, 50)
x = OI.normal((B, S, S, S, C_lat), 1)
tt = torch.full((B,), 37)
ctx = torch.arange(B) % 2
y = net([x.cuda(), tt] + ([ctx] if cond else []))
dt = torch.bfloat16 if _lib.precision() == "bf16" else torch.float16
tr, tr32 = {}, {}
with torch.no_grad():
    ref = ou.forward(P, x, tt, ctx=ctx if cond else None, emu=Emu(True, tr, dtype=dt))
    ref32 = ou.forward(P, x, tt, ctx=ctx if cond else None, emu=Emu(False, tr32))
for k, v in net.prog.outputs.items():
    if k in tr and v.numel() == tr[k].numel():
        p = v.float().cpu().reshape(tr[k].shape)
        e = ((p - tr[k]).norm() / tr[k].norm()).item()
        e32 = ((p - tr32[k]).norm() / tr32[k].norm()).item()
        ee = ((tr[k] - tr32[k]).norm() / tr32[k].norm()).item()
        print(f"{k:28s} prod-emu {e:.3e}  prod-fp32 {e32:.3e}  emu-fp32 {ee:.3e}")
print("final", ((y.cpu() - ref).norm() / ref.norm()).item(), ((y.cpu() - ref32).norm() / ref32.norm()).item(), "flag", _lib.debug_flag())
# attention internals of the first attention block, each stage checked against a CPU evaluation of the PRODUCT's own inputs
first = next((b["name"] for b in net.blocks if b.get("kind") == "attn"), None)
if first and not cond:
    O_ = net.prog.outputs
    g = lambda k: O_[k].float().cpu()
    nrm = g(f"{first}.norm")
    c = nrm.shape[-1]
    L = nrm[0].numel() // c
    f = nrm.reshape(B, L, c)
    r16 = lambda t: t.to(dt).float()
    for nm in ("query", "key", "value"):
        want = f @ r16(P[f"{first}.{nm}.kernel"]) + P[f"{first}.{nm}.bias"]
        got = g(f"{first}.{nm}")
        got = got.reshape(B, c, L).transpose(1, 2) if nm == "value" else got.reshape(B, L, c)
        print(f"{first}.{nm:6s} vs CPU dense of the product's norm: {((got - want).norm() / want.norm()).item():.3e}")
    q, k = g(f"{first}.query").reshape(B, L, c), g(f"{first}.key").reshape(B, L, c)
    v = g(f"{first}.value").reshape(B, c, L).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(1, 2) * c ** -0.5, -1) @ v
    got = g(f"{first}.flash").reshape(B, L, c)
    print(f"{first}.flash  vs CPU attention of the product's q,k,v: {((got - att).norm() / att.norm()).item():.3e}")
    want = nrm.reshape(B, L, c) + got @ r16(P[f"{first}.proj.kernel"]) + P[f"{first}.proj.bias"]
    print(f"{first}.proj   vs CPU: {((g(f'{first}.proj').reshape(B, L, c) - want).norm() / want.norm()).item():.3e}")
