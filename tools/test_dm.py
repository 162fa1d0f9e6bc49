#!/usr/bin/env python
"""`--test_dm` of the reference's entry points as a thin script over b200dm.

    reference  main.py:428-448 (flags :451-505)            python main.py --test_dm --suffix S --test_epoch E --vqvae_load_ckpt CKPT --timesteps T
               main_conditional_dm.py:195-215              python main_conditional_dm.py --test_dm ...   (conditional model)
    here       python tools/test_dm.py --test_dm [--conditional] --suffix S --test_epoch E --vqvae_load_ckpt CKPT --timesteps T
                      [--ckpt_root DIR] [--out_dir DIR] [--context 0|1] [--export_slices]

Same flag names and meaning; what the reference hard-codes is a default here: the checkpoint root
(/N/slate/aajais/checkpoints-dm -> --ckpt_root), the model geometry latent_size=int(64/4), num_embed=256, latent_channels=64.
`model.test(suffix)` then generates (10,16,16,16,64) latents, decodes them and writes `<out_dir>/<suffix>-<T>rsteps.npy`
(dm3d.py:534-545).  --export_slices additionally writes the mid-slice image the training callback logs
(WandbImageCallback.on_epoch_end, conditional_dm3d.py:24-58: images[:, :, :, slice_index, 0], first volume, grey-scale) for
context values 0 and 1, as .npy and .pgm (no matplotlib / wandb dependency).  Flags of the other modes (--train_*, --create_dataset,
--augment, ...) are accepted and rejected with a message: they are outside the sampling path.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    for flag in ("--create_dataset", "--augment", "--train_vq", "--train_dm", "--test_vq", "--save_best_only", "--test_run"):
        p.add_argument(flag, default=False, action="store_true", help="(reference flag; not on the sampling path)")
    p.add_argument("--test_dm", action="store_true", help="testing flag - Diffusion")
    p.add_argument("--dataset", type=str, default="both")
    p.add_argument("--exp_name", type=str, default="mask")
    p.add_argument("--lr", type=float, default=1e-4)
    p.add_argument("--lbs", type=int, default=5, help="Batch size per gpu")
    p.add_argument("--epochs", type=int, default=200)
    p.add_argument("--val_perc", type=float, default=0.1)
    p.add_argument("--suffix", default="basic", type=str, help="output or ckpts saved with this suffix")
    p.add_argument("--num_gpus", default=2, type=int)
    p.add_argument("--kernel_resize", action="store_true")
    p.add_argument("--test_epoch", type=int)
    p.add_argument("--vqvae_load_ckpt", type=str, default=None)
    p.add_argument("--timesteps", type=int, default=300)
    p.add_argument("--resume_ckpt", type=str)
    # what the reference hard-codes
    p.add_argument("--conditional", action="store_true", help="main_conditional_dm.py's model instead of main.py's")
    p.add_argument("--ckpt_root", default="/N/slate/aajais/checkpoints-dm")
    p.add_argument("--out_dir", default="./generated_images_dm3d")
    p.add_argument("--latent_size", type=int, default=int(64 / 4))
    p.add_argument("--num_embed", type=int, default=256)
    p.add_argument("--latent_channels", type=int, default=64)
    p.add_argument("--context", type=int, default=None, help="class id for the conditional model (0 healthy / 1 BraTS)")
    p.add_argument("--num_volumes", type=int, default=10)
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--export_slices", action="store_true", help="also write the callback's mid-slice images")
    p.add_argument("--random_init", action="store_true", help="skip load_weights (smoke runs without a checkpoint)")
    return p


def write_pgm(path, img):
    """8-bit grey-scale PGM of a 2-D array, min-max scaled like imshow(cmap='gray')."""
    a = np.asarray(img, dtype=np.float64)
    lo, hi = float(a.min()), float(a.max())
    u8 = np.zeros(a.shape, np.uint8) if hi <= lo else np.clip((a - lo) / (hi - lo) * 255.0 + 0.5, 0, 255).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (u8.shape[1], u8.shape[0]))
        f.write(u8.tobytes())


def export_slices(model, out_dir, tag, img_shape=None, slice_index=None, seed=None):
    """WandbImageCallback.on_epoch_end (conditional_dm3d.py:32-58) without wandb: for context 0 and 1, generate one volume over
    all timesteps, decode, take images[:, :, :, slice_index, 0] of the first volume."""
    S, C = model.latent_size, model.lc
    img_shape = img_shape or (1, S, S, S, C)
    paths = []
    os.makedirs(out_dir, exist_ok=True)
    for context_value in ((0, 1) if model.conditional else (None,)):
        kw = {} if context_value is None else dict(context_value=context_value)
        lat = model.generate(img_shape, last_step=0, seed=seed, **kw)
        images = model.vqvae_trainer.decoder(lat)
        si = slice_index if slice_index is not None else images.shape[3] // 2      # the reference's literal 64 = mid-slice of 128
        sl = images[:, :, :, si, 0].float().cpu().numpy()
        stem = os.path.join(out_dir, f"{tag}-slice-ctx{0 if context_value is None else context_value}")
        np.save(stem + ".npy", sl)
        write_pgm(stem + ".pgm", sl[0])
        paths.append(stem)
    return paths


def run(args):
    if not args.test_dm:
        raise SystemExit("tools/test_dm.py implements --test_dm only (the sampling + decode path); training / dataset modes are out of scope")
    import b200dm
    print(f"Testing Diffusion Model with ckpt - {args.suffix}-{args.test_epoch}")
    cls = b200dm.ConditionalDiffusionModel if args.conditional else b200dm.DiffusionModel
    model = cls(latent_size=args.latent_size, num_embed=args.num_embed, latent_channels=args.latent_channels,
                vqvae_load_ckpt=args.vqvae_load_ckpt, args=types.SimpleNamespace(timesteps=args.timesteps, num_gpus=args.num_gpus,
                                                                                kernel_resize=args.kernel_resize, bs=args.lbs * args.num_gpus))
    if not args.random_init:
        model.load_weights(os.path.join(args.ckpt_root, args.suffix, str(args.test_epoch) + ".ckpt"))
    suffix = args.suffix + "epoch" + str(args.test_epoch)
    S, C = args.latent_size, args.latent_channels
    kw = dict(context=args.context) if args.conditional and args.context is not None else {}
    images = model.test(suffix, shape=(args.num_volumes, S, S, S, C), out_dir=args.out_dir, seed=args.seed, **kw)
    print(f"wrote {os.path.join(args.out_dir, suffix)}-{args.timesteps}rsteps.npy {tuple(images.shape)}")
    if args.export_slices:
        for p in export_slices(model, args.out_dir, suffix, seed=args.seed):
            print("wrote", p + ".npy/.pgm")
    return images


if __name__ == "__main__":
    run(build_parser().parse_args())
