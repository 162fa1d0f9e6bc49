"""tests/golden/chain1000_cfg1.npz: the FULL 1000-step DDPM chain of BASELINE config 1 (unconditional dm3d U-Net, 16^3 x 8 latent,
batch 1) evaluated by the fp32 CPU oracle: x_T = oracle N(0,1) seed 1234, per-step noise = oracle Philox4x32-10 stream (seed 1234,
sample id 0) -- the stream the CUDA update kernel draws in-register, so the production graph path can be compared end to end.
Stress-initialised weights seed 0 (no numerically dead branch).  Also records the latents after 10 / 100 / 500 steps.
Not a reference (TensorFlow) output: parity unpinned (oracle/__init__.py).  ~3 min on 8 cores:  python tools/make_golden_chain1000.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import init as OI, sampler as OS
from oracle.schedule import Betas
from oracle.unet import UNet

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "chain1000_cfg1.npz")
S, C, T, seed = 16, 8, 1000, 1234
u = UNet(S, C, [64, 128, 256], [False, False, True, True], first_conv_channels=64)
P = OI.make_params(u.spec(), 0, "stress")
shape = (1, S, S, S, C)
xT = OI.normal(shape, seed)
snaps = {}


def rec(i, x, eps):
    if i in (T - 10, T - 100, T - 500):
        snaps[f"x_before_t{i}"] = x.numpy().astype(np.float32).copy()


t0 = time.time()
with torch.no_grad():
    lat = OS.generate(lambda z, i: u.forward(P, z, torch.full((1,), i)), Betas(T), xT, seed=seed, sample_ids=np.arange(1), record=rec)
np.savez_compressed(out, x_T=xT.numpy(), latents=lat.numpy().astype(np.float32), seed=seed, **snaps)
print(f"wrote {out} in {time.time() - t0:.0f} s; latents std {lat.std().item():.4f}, size {os.path.getsize(out)}")
