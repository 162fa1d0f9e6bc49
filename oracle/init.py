"""Seeded synthetic weights and inputs (oracle; test infrastructure only).

Initialiser distributions follow the reference:
  'vs1'    VarianceScaling(1.0, fan_avg, uniform)  (kernel_init(1.0), dm3d.py:11-15) == glorot_uniform
  'vs0'    VarianceScaling(1e-10, ...)             (kernel_init(0.0): ~zero last convs)
  'glorot' Keras default glorot_uniform
  'embed'  Keras Embedding default uniform(-0.05, 0.05)
``mode='keras'`` reproduces those; ``mode='stress'`` makes every branch numerically alive
(every kernel VarianceScaling(1.0); BN gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5);
LN/GN gamma~U(.5,1.5), beta~N(0,.1); biases ~N(0,.05); PReLU alpha~U(0,.3)).
"""
from __future__ import annotations

import numpy as np
import torch


def _fans(shape):
    if len(shape) == 2:
        return shape[0], shape[1]
    rec = int(np.prod(shape[:-2]))
    return shape[-2] * rec, shape[-1] * rec


def make_params(spec, seed=0, mode="stress", dtype=torch.float32):
    rng = np.random.default_rng(seed)
    P = {}
    for name, shape, init in spec:
        leaf = name.rsplit(".", 1)[-1]
        if init in ("vs1", "vs0", "glorot"):
            fi, fo = _fans(shape)
            scale = 1.0 if (init != "vs0" or mode == "stress") else 1e-10
            lim = np.sqrt(3.0 * scale / ((fi + fo) / 2.0))
            a = rng.uniform(-lim, lim, size=shape)
        elif init == "embed":
            a = rng.uniform(-0.05, 0.05, size=shape) * (10.0 if mode == "stress" else 1.0)
        elif mode == "keras":
            a = np.ones(shape) if init == "ones" else np.zeros(shape)
        elif leaf == "gamma":
            a = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == "beta":
            a = rng.normal(0, 0.1, size=shape)
        elif leaf == "mean":
            a = rng.normal(0, 0.1, size=shape)
        elif leaf == "var":
            a = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == "alpha":
            a = rng.uniform(0.0, 0.3, size=shape)
        elif leaf == "bias":
            a = rng.normal(0, 0.05, size=shape)
        else:
            raise ValueError(f"no stress initialiser for {name}")
        P[name] = torch.tensor(np.asarray(a, dtype=np.float32)).to(dtype)
    return P


def codebook(K, D, layout="DK", seed=3, kind="uniform"):
    """kind 'uniform': random_uniform(-0.05,0.05) (vqgan_attn_cp.py:154); 'he': HeUniform on (D,K) (vqvae3d_monai.py:123)."""
    rng = np.random.default_rng(seed)
    if kind == "he":
        lim = np.sqrt(6.0 / D)
        e = rng.uniform(-lim, lim, size=(D, K))
        e = e if layout == "DK" else e.T
    else:
        e = rng.uniform(-0.05, 0.05, size=(K, D))
        e = e.T if layout == "DK" else e
    return torch.tensor(np.ascontiguousarray(e, dtype=np.float32))


def normal(shape, seed, scale=1.0):
    return torch.tensor((np.random.default_rng(seed).standard_normal(size=shape) * scale).astype(np.float32))
