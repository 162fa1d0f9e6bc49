"""Full cfg-2 chain through the public API: 1000 DDPM steps, batch 8, 32^3 x 256 latent (GPU box): time, finiteness, watchdog flag,
bit-reproducibility of the Philox-driven chain."""
import os, sys, types, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200dm
from b200dm import _lib as L
S, C, B, T = 32, 256, 8, 1000
dm = b200dm.ConditionalDiffusionModel(S, 1024, C, None, types.SimpleNamespace(timesteps=T, num_gpus=1, kernel_resize=False, bs=B))
shape = (B, S, S, S, C)
ctx = torch.arange(B) % 2
lat = dm.generate(shape, last_step=T - 3, seed=1, context=ctx)   # warm
torch.cuda.synchronize(); t0 = time.perf_counter()
lat = dm.generate(shape, seed=1, context=ctx)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
lat2 = dm.generate(shape, seed=1, context=ctx)
print("1000-step chain: %.3f s, %.1f volume-steps/s, finite=%s, flag=%d, reproducible=%s, |lat|max=%.3f" %
      (dt, B * T / dt, bool(torch.isfinite(lat).all()), L.debug_flag(), bool(torch.equal(lat, lat2)), lat.abs().max().item()))
