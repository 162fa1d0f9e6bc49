"""Watchdog stress: replay each hot path many times and report the tcgen05 / TMA watchdog flag (0 = clean) -- the flag is set by any
bounded mbarrier wait that times out (code = which one).   python tools/stress_flags.py [cfg3 N] [cfg4 N] [cfg2 N]"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dm
from b200dm import _lib as L

args = sys.argv[1:]
want = {args[i]: int(args[i + 1]) for i in range(0, len(args), 2)} or {"cfg3": 100, "cfg4": 30, "cfg2": 1500}
dev = torch.device("cuda", 0)


def report(name, t0, extra=""):
    torch.cuda.synchronize()
    print(f"{name}: flag {L.debug_flag():#x} {extra} {time.time() - t0:.1f}s", flush=True)


if "cfg3" in want:
    vq = b200dm.VQGAN(num_channels=(32, 64, 128), num_embeddings=1024, embedding_dim=256)
    z = torch.randn(16, 32, 32, 32, 256, device=dev) * 0.05
    q, _, _ = vq.quantizer.quantize(z)
    vol = vq.decoder(q)
    report("cfg3 first decode", time.time(), f"finite {bool(torch.isfinite(vol).all())}")
    t0 = time.time()
    for i in range(want["cfg3"]):
        vq.decoder.prog.run()
        if i % 25 == 24:
            report(f"cfg3 decode x{i + 1}", t0)
    del vq, z, q, vol
if "cfg4" in want:
    B, S, C, T = 1, 32, 256, 1000
    dm = b200dm.DiffusionModel(S, 256, C, None, types.SimpleNamespace(timesteps=T, num_gpus=1, kernel_resize=False, bs=B))
    dm.network = b200dm.build_model(S, C, [64, 128, 256], [True, False, True])
    seq = list(range(T - 1, T - 1 - 4 * 16, -4))
    t0 = time.time()
    for i in range(want["cfg4"]):
        lat = dm.generate((B, S, S, S, C), seed=1, sampler="ddim", timestep_seq=seq)
        if i % 10 == 9 or i == 0:
            report(f"cfg4 generate(16 ddim steps) x{i + 1}", t0, f"finite {bool(torch.isfinite(lat).all())} fused {dm._step['fused']}")
    del dm
if "cfg2" in want:
    B, S, C, T = 8, 32, 256, 1000
    dm = b200dm.ConditionalDiffusionModel(S, 1024, C, None, types.SimpleNamespace(timesteps=T, num_gpus=1, kernel_resize=False, bs=B))
    t0 = time.time()
    n = want["cfg2"]
    lat = dm.generate((B, S, S, S, C), seed=1, last_step=T - min(n, T), context=torch.arange(B) % 2)
    report(f"cfg2 generate({min(n, T)} ddpm steps)", t0, f"finite {bool(torch.isfinite(lat).all())} fused {dm._step['fused']}")
