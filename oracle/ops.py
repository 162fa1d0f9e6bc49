"""Keras-2 op semantics restated on PyTorch-CPU (oracle; test infrastructure only).

All tensors at this API are channels-last (N, D, H, W, C) like the reference's.
Weight layouts are the Keras ones:
  Conv3D kernel (kd,kh,kw,Cin,Cout); Conv3DTranspose kernel (kd,kh,kw,Cout,Cin);
  Dense kernel (in,out); BN gamma/beta/moving_mean/moving_variance (C); PReLU alpha (d,h,w,C).
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F


class Emu:
    """Rounding policy.  ``Emu(False)`` = pure reference arithmetic (no rounding).
    ``Emu(True)`` rounds to bf16 at the CUDA path's bf16 storage points (activations
    written to HBM, weights fed to the tensor cores)."""

    def __init__(self, bf16: bool = False, trace: dict | None = None, dtype: torch.dtype = torch.bfloat16):
        self.bf16 = bf16
        self.trace = trace  # optional {tag: tensor} of every tagged storage point (layer-by-layer debugging)
        self.dtype = dtype  # the CUDA build's 16-bit storage type: torch.bfloat16 (libb200dm.so) or torch.float16 (libb200dm_f16.so)

    def a(self, x: torch.Tensor, tag: str | None = None) -> torch.Tensor:  # activation storage point
        y = x.to(self.dtype).to(x.dtype) if self.bf16 else x
        if self.trace is not None and tag is not None:
            self.trace[tag] = y
        return y

    def w(self, x: torch.Tensor) -> torch.Tensor:  # tensor-core weight operand
        return x.to(self.dtype).to(x.dtype) if self.bf16 else x


EXACT = Emu(False)


def _cl_to_cf(x):  # NDHWC -> NCDHW
    return x.permute(0, 4, 1, 2, 3)


def _cf_to_cl(x):
    return x.permute(0, 2, 3, 4, 1)


def same_pads(in_size: int, k: int, s: int):
    """TF/Keras 'same' padding for one axis: out=ceil(in/s); total=max((out-1)s+k-in,0);
    before=total//2, after=rest (asymmetric: k=3,s=2,even in -> (0,1); k=4,s=1 -> (1,2))."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


def conv3d(x, kernel, bias=None, stride: int = 1, padding: str = "same"):
    """keras.layers.Conv3D (cross-correlation, channels-last).  kernel (kd,kh,kw,Cin,Cout)."""
    k = kernel.shape[0]
    w = kernel.permute(4, 3, 0, 1, 2).contiguous()
    xc = _cl_to_cf(x)
    if padding == "same":
        pads = []
        for ax in (3, 2, 1):  # F.pad wants last dim first: W, H, D
            b, a = same_pads(x.shape[ax], k, stride)
            pads += [b, a]
        xc = F.pad(xc, pads)
    y = F.conv3d(xc, w, bias, stride=stride)
    return _cf_to_cl(y)


def conv3d_transpose(x, kernel, bias=None):
    """keras.layers.Conv3DTranspose(k=4, strides=2, padding='same'); kernel (kd,kh,kw,Cout,Cin).
    out[o] = sum_i in[i] * w[o + 1 - 2 i]  ==  torch conv_transpose3d(stride=2, padding=1)
    with weight (Cin, Cout, kd,kh,kw) = kernel.permute(4,3,0,1,2), no flip."""
    assert kernel.shape[0] == 4
    w = kernel.permute(4, 3, 0, 1, 2).contiguous()
    y = F.conv_transpose3d(_cl_to_cf(x), w, bias, stride=2, padding=1)
    return _cf_to_cl(y)


def upsample_nearest2(x):
    """keras.layers.UpSampling3D(size=2): repeat each voxel 2x along D,H,W."""
    return x.repeat_interleave(2, 1).repeat_interleave(2, 2).repeat_interleave(2, 3)


def dense(x, kernel, bias=None):
    y = x @ kernel
    return y if bias is None else y + bias


def swish(x):
    return x * torch.sigmoid(x)


def batchnorm_infer(x, gamma, beta, mean, var, eps: float = 1e-3):
    """keras BatchNormalization, training=False: gamma*(x-mean)/sqrt(var+eps)+beta.
    Keras default epsilon is 1e-3."""
    inv = gamma * torch.rsqrt(var + eps)
    return x * inv + (beta - mean * inv)


def groupnorm(x, gamma, beta, groups: int, eps: float):
    """keras GroupNormalization(axis=-1): per (sample, group) mean / biased variance over
    (D,H,W,C/G); groups are contiguous channel blocks."""
    n, d, h, w, c = x.shape
    xg = x.reshape(n, d * h * w, groups, c // groups)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    var = xg.var(dim=(1, 3), unbiased=False, keepdim=True)
    y = (xg - mean) * torch.rsqrt(var + eps)
    return y.reshape(n, d, h, w, c) * gamma + beta


def layernorm(x, gamma, beta, eps: float = 1e-3):
    """keras LayerNormalization(axis=-1), default epsilon 1e-3."""
    mean = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def prelu(x, alpha):
    """keras PReLU(): alpha has one entry per non-batch element (d,h,w,c)."""
    return torch.clamp(x, min=0) + alpha * torch.clamp(x, max=0)


def attention_core(q, k, v, scale: float, emu: Emu = EXACT):
    """softmax(q k^T * scale) v over flattened voxel tokens; q (B,Lq,C), k/v (B,Lk,C).
    Scores and softmax in full precision; probabilities are a bf16 storage point."""
    s = torch.einsum("blc,bLc->blL", q, k) * scale
    p = emu.a(torch.softmax(s, dim=-1))
    return torch.einsum("blL,bLc->blc", p, v)


def time_embedding(t, dim: int):
    """TimeEmbedding (dm3d.py:177-191): f_j = exp(-j ln(10000)/(half-1)); [sin(t f), cos(t f)]."""
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -emb)
    e = t.to(torch.float32)[:, None] * freqs[None, :]
    return torch.cat([torch.sin(e), torch.cos(e)], dim=-1)
