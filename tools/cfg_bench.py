"""Timing of the non-headline BASELINE.json configs on one GPU (CUDA events; synthetic inputs, random-init weights).

    python tools/cfg_bench.py cfg1|cfg3|cfg3-monai|cfg4 [--batch B] [--steps K]

cfg1  unconditional dm3d U-Net, 50-step DDPM, batch 1, 16^3 x 8 latent        -> volume-steps/s
cfg3  vqgan_attn_cp quantize (K=1024, D=256) + Decoder (32,64,128): 32^3 -> 128^3, batch 16 -> volumes/s, VQ rows/s
cfg4  dm3d U-Net with self-attention at the highest latent resolution (32^3 = 32768 tokens, d=64), 250-step DDIM
Each prints one JSON line (these are parity-test configurations, not the bench.py headline)."""
import argparse, json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200dm
from b200dm import _lib


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def args_ns(T, B):
    return types.SimpleNamespace(timesteps=T, num_gpus=1, kernel_resize=False, bs=B)


def cfg1(a):
    B, S, C, T = a.batch or 1, 16, 8, 50
    dm = b200dm.DiffusionModel(S, 256, C, None, args_ns(T, B))
    shape = (B, S, S, S, C)
    lat = dm.generate(shape, seed=1)                      # compiles + captures
    assert torch.isfinite(lat).all() and _lib.debug_flag() == 0
    ms = timed(lambda: dm.generate(shape, seed=1), 3, 1)
    return {"config": "cfg-1 unconditional dm3d U-Net, 50-step DDPM, 16^3x8 latent", "batch": B, "ms_per_chain": ms,
            "ms_per_step": ms / T, "volume_steps_per_s": B * T / (ms * 1e-3), "launches_per_step": dm._step["nets"][0].prog.num_launches + 2}


def cfg4(a):
    B, S, C, T, steps = a.batch or 1, 32, 256, 1000, 250
    dm = b200dm.DiffusionModel(S, 256, C, None, args_ns(T, B))
    dm.network = b200dm.build_model(S, C, [64, 128, 256], [True, False, True])
    shape = (B, S, S, S, C)
    K = a.steps or 8
    lat = dm.generate(shape, seed=1, sampler="ddim", steps=steps, last_step=T - 1 - 4 * 3)    # compile; a few DDIM steps
    assert torch.isfinite(lat).all() and _lib.debug_flag() == 0
    st = dm._step
    net = st["nets"][0]
    ms = timed(lambda: dm._run_step_eager(st), K, 2)
    rows = net.prog.run_timed()
    attn = [r for r in rows if r[0] == "attn"]
    a_ms, a_fl = sum(r[3] for r in attn), sum(r[2] for r in attn)
    tot = sum(r[3] for r in rows)
    return {"config": "cfg-4 dm3d U-Net, self-attention at 32^3 (L=32768, d=64), DDIM (250 of 1000 steps)", "batch": B,
            "ms_per_step": ms, "volume_steps_per_s": B / (ms * 1e-3), "seconds_per_250_step_chain": 250 * ms * 1e-3,
            "flash_attention": {"launches": len(attn), "ms": a_ms, "tflops": a_fl / a_ms / 1e9, "share_of_step": a_ms / tot}}


def cfg3(a, family="attn_cp"):
    B = a.batch or 16
    dev = torch.device("cuda", 0)
    if family == "attn_cp":
        vq = b200dm.VQGAN(num_channels=(32, 64, 128), num_embeddings=1024, embedding_dim=256)
        s = 32
    else:  # the DM's actual first stage (dm3d.py:386-404): 8^3 x 256 -> 128^3, channels (256,128,64,32), R=5
        vq = b200dm.VQVAE(in_channels=1, out_channels=1, num_channels=(32, 64, 128, 256), num_res_channels=(32, 64, 128, 256),
                          num_res_layers=5, num_embeddings=1024, embedding_dim=256, latent_size=8)
        s = 8
    D = 256
    z = torch.randn(B, s, s, s, D, device=dev) * 0.05
    q, idx, perp = vq.quantizer.quantize(z)
    vol = vq.decoder(q)
    assert torch.isfinite(vol).all() and _lib.debug_flag() == 0 and tuple(vol.shape) == (B, 128, 128, 128, 1)
    ms_q = timed(lambda: vq.quantizer.quantize(z), 3, 1)
    ms_d = timed(lambda: vq.decoder.prog.run(), 3, 1)
    rows = vq.decoder.prog.run_timed()
    kinds = {}
    for r in rows:
        e = kinds.setdefault(r[0], [0, 0.0, 0.0]); e[0] += 1; e[1] += r[3]; e[2] += r[2]
    N = B * s ** 3
    return {"config": f"cfg-3 {family}: VQ quantize (K=1024, D=256) + decoder {s}^3 -> 128^3", "batch": B,
            "quantize_ms": ms_q, "vq_rows_per_s": N / (ms_q * 1e-3), "vq_tflops_fp32": 2.0 * N * 1024 * D / ms_q / 1e9,
            "decode_ms": ms_d, "decode_volumes_per_s": B / (ms_d * 1e-3), "e2e_volumes_per_s": B / ((ms_q + ms_d) * 1e-3),
            "decoder_ops": {k: {"launches": v[0], "ms": v[1], ("tflops" if k in ("conv", "conv_halo") else "gbps"):
                                (v[2] / v[1] / 1e9 if k in ("conv", "conv_halo") else v[2] / v[1] / 1e6)} for k, v in kinds.items()}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cfg")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=None)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    out = {"cfg1": cfg1, "cfg4": cfg4, "cfg3": cfg3, "cfg3-monai": lambda a: cfg3(a, "monai")}[a.cfg](a)
    print(json.dumps(out), flush=True)
