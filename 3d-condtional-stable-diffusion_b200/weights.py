"""Weights: canonical names, Keras layouts, initialisers, npz I/O.

The reference saves TF object-graph checkpoints (``ModelCheckpoint(save_weights_only=True)``,
main.py:392-396) whose keys are object-graph paths, not layer names.  TensorFlow is not available in this
environment, so the contract here is: one canonical structural name per tensor
(``down.0.res.1.conv2.kernel``), tensors kept in the reference's Keras layouts

    Conv3D kernel (kd,kh,kw,Cin,Cout) | Conv3DTranspose kernel (kd,kh,kw,Cout,Cin) | Dense kernel (in,out)
    BN gamma/beta/mean/var (C) | LN/GN gamma/beta (C) | PReLU alpha (d,h,w,C) | Embedding (vocab,dim)
    codebook (D,K) [monai, vqgan] or (K,D) [gnorm, stride, attn_cp]

and a flat ``{name: ndarray}`` .npz as the interchange file.  ``spec`` lists tensors in the order the reference
code constructs its layers, so a TF-side exporter can zip ``model.network.weights`` onto it with shape checks
(tools/tf_export_npz.py documents that script; it cannot run here).
"""
from __future__ import annotations

import numpy as np
import torch


def _fans(shape):
    if len(shape) == 2:
        return shape[0], shape[1]
    rec = int(np.prod(shape[:-2]))
    return shape[-2] * rec, shape[-1] * rec


def init_params(spec, seed=0, mode="keras"):
    """Random-init weights with the reference's initialiser distributions (dm3d.py:11-15: VarianceScaling(scale,
    fan_avg, uniform), scale=max(s,1e-10); Keras defaults glorot_uniform / zeros / ones / uniform(-.05,.05)).
    mode='stress' replaces the ~zero / constant initialisers by O(1) random values so no branch is numerically dead."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape, init in spec:
        leaf = name.rsplit(".", 1)[-1]
        if init in ("vs1", "vs0", "glorot"):
            fi, fo = _fans(shape)
            scale = 1e-10 if (init == "vs0" and mode == "keras") else 1.0
            lim = np.sqrt(3.0 * scale / ((fi + fo) / 2.0))
            a = rng.uniform(-lim, lim, size=shape)
        elif init == "embed":
            a = rng.uniform(-0.05, 0.05, size=shape) * (10.0 if mode == "stress" else 1.0)
        elif mode == "keras":
            a = np.ones(shape) if init == "ones" else np.zeros(shape)
        elif leaf == "gamma":
            a = rng.uniform(0.5, 1.5, size=shape)
        elif leaf in ("beta", "mean"):
            a = rng.normal(0, 0.1, size=shape)
        elif leaf == "var":
            a = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == "alpha":
            a = rng.uniform(0.0, 0.3, size=shape)
        elif leaf == "bias":
            a = rng.normal(0, 0.05, size=shape)
        else:
            raise ValueError(f"no initialiser for {name}")
        out[name] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    return out


def check_against_spec(params, spec):
    names = {n for n, _, _ in spec}
    missing = [n for n in names if n not in params]
    extra = [n for n in params if n not in names]
    if missing or extra:
        raise KeyError(f"weights do not match the model: missing {missing[:5]} ({len(missing)}), unexpected {extra[:5]} ({len(extra)})")
    for n, shape, _ in spec:
        if tuple(params[n].shape) != tuple(shape):
            raise ValueError(f"{n}: expected Keras-layout shape {tuple(shape)}, got {tuple(params[n].shape)}")


def save_npz(path, params):
    np.savez(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in params.items()})


def load_npz(path, prefix=""):
    with np.load(path) as z:
        return {k[len(prefix):]: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if k.startswith(prefix)}
