"""Generate tests/golden/*.npz from the oracle (seeded).  The reference (TensorFlow) cannot run here, so these vectors
pin the ORACLE against regressions and give the GPU tests fixed targets; they are not reference outputs (parity unpinned,
see oracle/__init__.py).  Re-run:  python tools/make_golden.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import init as OI, ops as O, sampler as OS, first_stage as OF, philox
from oracle.schedule import Betas
from oracle.unet import UNet

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(out, exist_ok=True)
torch.set_num_threads(4)

# 1. unconditional U-Net forward, 8^3 x 8, B=2, stress weights seed 0
u = UNet(8, 8, [64, 128, 256], [False, False, True, True], first_conv_channels=64)
P = OI.make_params(u.spec(), 0, "stress")
x = OI.normal((2, 8, 8, 8, 8), 1)
eps = u.forward(P, x, torch.tensor([37, 37]))
np.savez_compressed(os.path.join(out, "unet_uncond_8.npz"), x=x.numpy(), t=37, eps=eps.numpy().astype(np.float32))

# 2. conditional U-Net forward, 8^3 x 16, B=2
uc = UNet(8, 16, [64, 128, 256], [False, False, True, True], first_conv_channels=32, conditional=True)
Pc = OI.make_params(uc.spec(), 0, "stress")
xc = OI.normal((2, 8, 8, 8, 16), 1)
epsc = uc.forward(Pc, xc, torch.tensor([21, 21]), ctx=torch.tensor([0, 1]))
np.savez_compressed(os.path.join(out, "unet_cond_8.npz"), x=xc.numpy(), t=21, ctx=np.array([0, 1]), eps=epsc.numpy().astype(np.float32))

# 3. DDPM chain T=12 with injected noise (uses model 1)
T = 12
b = Betas(T)
shape = (2, 8, 8, 8, 8)
xT = OI.normal(shape, 1234)
noises = {i: OI.normal(shape, 100 + i) for i in range(1, T)}
lat = OS.generate(lambda z, i: u.forward(P, z, torch.full((2,), i)), b, xT, noises=noises)
np.savez_compressed(os.path.join(out, "chain_T12.npz"), latents=lat.numpy().astype(np.float32))

# 4. VQ indices + decoders
cb = OI.codebook(256, 64, "KD", seed=3)
z = OI.normal((2, 4, 4, 4, 64), 11, 0.05)
q, idx, perp, counts = OF.quantize(z, cb, "KD")
dec = OF.AttnCpDecoder(64, 1, (32, 64, 128))
Pd = OI.make_params(dec.spec(), 5, "stress")
vol = dec.forward(Pd, q)
md = OF.MonaiDecoder(64, 1, (32, 64), 2, (32, 64), 4)
Pm = OI.make_params(md.spec(), 5, "stress")
volm = md.forward(Pm, z)
np.savez_compressed(os.path.join(out, "vq_decode.npz"), idx=idx.numpy(), perplexity=float(perp), vol_attn_cp=vol.numpy().astype(np.float32),
                    vol_monai=volm.numpy().astype(np.float32))

# 5. Philox normals + one update step
zz = philox.normal(1234, 17, [5, 6, 7], 4 * 4 * 4 * 8)
np.savez_compressed(os.path.join(out, "philox.npz"), z=zz)
print("golden vectors written to", out, {f: os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)})
