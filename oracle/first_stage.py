"""First-stage quantizer and 3D decoders (oracle; test infrastructure only).

Restates
  VectorQuantizer.get_code_indices / call   networks/vqvae3d_monai.py:165-177,133-163
      (variants vqgan.py:204-216, vqgan_gnorm.py:203-215, vqgan_stride.py:204-216,
       vqgan_attn_cp.py:189-201,203-247); codebook stored (D,K) or (K,D) per variant
  D1  vqvae3d_monai.Decoder :309-391 + VQVAEResidualUnit :218-234
  D5  vqgan_attn_cp.Decoder :339-427 + VQVAEResidualUnit :250-276
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .ops import Emu, EXACT


# ------------------------------------------------------------------ quantizer
def code_distances(flat: torch.Tensor, codebook: torch.Tensor, layout: str = "DK"):
    """||x||^2 + ||e||^2 - 2 x.e  (N,K); codebook (D,K) for layout 'DK', (K,D) for 'KD'."""
    E = codebook if layout == "DK" else codebook.t()
    sim = flat @ E
    return (flat ** 2).sum(1, keepdim=True) + (E ** 2).sum(0) - 2 * sim


def get_code_indices(flat, codebook, layout="DK"):
    """argmin over codes, lowest index on ties (tf.argmin)."""
    d = code_distances(flat, codebook, layout)
    # torch.argmin does not promise first-min on ties: make it explicit
    m = d.min(dim=1, keepdim=True).values
    K = d.shape[1]
    idx = torch.where(d == m, torch.arange(K)[None, :], torch.full((1, 1), K)).min(dim=1).values
    return idx.to(torch.int64)


def get_code_indices_exact(flat, codebook, layout="DK"):
    """Same argmin evaluated in float64 on the float32 inputs (rounding-free ranking) and the
    float64 margin between best and second-best distance, to classify near-ties."""
    d = code_distances(flat.double(), codebook.double(), layout)
    top2 = torch.topk(d, 2, dim=1, largest=False).values
    m = top2[:, :1]
    K = d.shape[1]
    idx = torch.where(d == m, torch.arange(K)[None, :], torch.full((1, 1), K)).min(dim=1).values
    return idx.to(torch.int64), (top2[:, 1] - top2[:, 0])


def quantize(x, codebook, layout="DK"):
    """VectorQuantizer.call at inference: -> (quantized (same shape), indices, perplexity, counts).
    Returns the gathered code rows q (the reference's STE form x+(q-x) equals q to 1 ulp)."""
    D = x.shape[-1]
    flat = x.reshape(-1, D)
    idx = get_code_indices(flat, codebook, layout)
    E = codebook.t() if layout == "DK" else codebook  # (K,D)
    q = E[idx].reshape(x.shape)
    K = E.shape[0]
    counts = torch.bincount(idx, minlength=K)
    p = counts.double() / flat.shape[0]
    perplexity = torch.exp(-(p * torch.log(p + 1e-10)).sum()).float()
    return q, idx, perplexity, counts


# ------------------------------------------------------------------ decoder D1 (monai)
class MonaiDecoder:
    """vqvae3d_monai.Decoder: Conv3(D->c_top) PReLU; per level R x ResUnit, ConvT(k4,s2) (+ReLU if not last)."""

    def __init__(self, in_channels, out_channels, num_channels, num_res_layers, num_res_channels, in_size):
        self.cin, self.cout = in_channels, out_channels
        self.ch = list(reversed(num_channels))
        self.rch = list(reversed(num_res_channels))
        self.R, self.s0 = num_res_layers, in_size

    def spec(self):
        sp, s = [], self.s0
        c = self.ch[0]
        sp += [("stem.kernel", (3, 3, 3, self.cin, c), "glorot"), ("stem.bias", (c,), "zeros"),
               ("stem.prelu.alpha", (s, s, s, c), "zeros")]
        for i, c in enumerate(self.ch):
            for j in range(self.R):
                n = f"level.{i}.res.{j}"
                rc = self.rch[i]
                sp += [(f"{n}.conv1.kernel", (3, 3, 3, c, rc), "glorot"), (f"{n}.conv1.bias", (rc,), "zeros"),
                       (f"{n}.conv2.kernel", (3, 3, 3, rc, c), "glorot"), (f"{n}.conv2.bias", (c,), "zeros"),
                       (f"{n}.norm.gamma", (c,), "ones"), (f"{n}.norm.beta", (c,), "zeros"),
                       (f"{n}.norm.mean", (c,), "zeros"), (f"{n}.norm.var", (c,), "ones"),
                       (f"{n}.prelu.alpha", (s, s, s, c), "zeros")]
            out = self.cout if i == len(self.ch) - 1 else self.ch[i + 1]
            sp += [(f"level.{i}.up.kernel", (4, 4, 4, out, c), "glorot"), (f"level.{i}.up.bias", (out,), "zeros")]
            s *= 2
        return sp

    def forward(self, P, z, emu: Emu = EXACT):
        x = emu.a(ops.prelu(ops.conv3d(emu.a(z), emu.w(P["stem.kernel"]), P["stem.bias"]), P["stem.prelu.alpha"]))
        for i in range(len(self.ch)):
            for j in range(self.R):
                n = f"level.{i}.res.{j}"
                h = emu.a(torch.relu(ops.conv3d(x, emu.w(P[f"{n}.conv1.kernel"]), P[f"{n}.conv1.bias"])))
                h = ops.conv3d(h, emu.w(P[f"{n}.conv2.kernel"]), P[f"{n}.conv2.bias"])
                h = ops.batchnorm_infer(h, P[f"{n}.norm.gamma"], P[f"{n}.norm.beta"], P[f"{n}.norm.mean"], P[f"{n}.norm.var"])
                x = emu.a(torch.relu(x + ops.prelu(h, P[f"{n}.prelu.alpha"])))
            x = ops.conv3d_transpose(x, emu.w(P[f"level.{i}.up.kernel"]), P[f"level.{i}.up.bias"])
            if i != len(self.ch) - 1:
                x = emu.a(torch.relu(x))
        return x


# ------------------------------------------------------------------ encoder (monai)
class MonaiEncoder:
    """vqvae3d_monai.Encoder (:237-306): per level Conv3D(c_i, k=4, strides=2, 'same') ReLU, R x VQVAEResidualUnit (:218-234);
    then Conv3D(embedding_dim, 3, 'same') PReLU.  Dropout is an inference no-op.  PReLU alphas are per voxel (d,h,w,c)."""

    def __init__(self, in_channels, out_channels, num_channels, num_res_layers, num_res_channels, in_size):
        self.cin, self.cout, self.ch, self.rch = in_channels, out_channels, list(num_channels), list(num_res_channels)
        self.R, self.s0 = num_res_layers, in_size

    def spec(self):
        sp, s, cin = [], self.s0, self.cin
        for i, c in enumerate(self.ch):
            s //= 2
            sp += [(f"level.{i}.down.kernel", (4, 4, 4, cin, c), "glorot"), (f"level.{i}.down.bias", (c,), "zeros")]
            for j in range(self.R):
                n, rc = f"level.{i}.res.{j}", self.rch[i]
                sp += [(f"{n}.conv1.kernel", (3, 3, 3, c, rc), "glorot"), (f"{n}.conv1.bias", (rc,), "zeros"),
                       (f"{n}.conv2.kernel", (3, 3, 3, rc, c), "glorot"), (f"{n}.conv2.bias", (c,), "zeros"),
                       (f"{n}.norm.gamma", (c,), "ones"), (f"{n}.norm.beta", (c,), "zeros"),
                       (f"{n}.norm.mean", (c,), "zeros"), (f"{n}.norm.var", (c,), "ones"),
                       (f"{n}.prelu.alpha", (s, s, s, c), "zeros")]
            cin = c
        sp += [("head.kernel", (3, 3, 3, cin, self.cout), "glorot"), ("head.bias", (self.cout,), "zeros"),
               ("head.prelu.alpha", (s, s, s, self.cout), "zeros")]
        return sp

    def forward(self, P, x, emu: Emu = EXACT):
        x = emu.a(x)
        for i in range(len(self.ch)):
            x = emu.a(torch.relu(ops.conv3d(x, emu.w(P[f"level.{i}.down.kernel"]), P[f"level.{i}.down.bias"], stride=2)))
            for j in range(self.R):
                n = f"level.{i}.res.{j}"
                h = emu.a(torch.relu(ops.conv3d(x, emu.w(P[f"{n}.conv1.kernel"]), P[f"{n}.conv1.bias"])))
                h = ops.conv3d(h, emu.w(P[f"{n}.conv2.kernel"]), P[f"{n}.conv2.bias"])
                h = ops.batchnorm_infer(h, P[f"{n}.norm.gamma"], P[f"{n}.norm.beta"], P[f"{n}.norm.mean"], P[f"{n}.norm.var"])
                x = emu.a(torch.relu(x + ops.prelu(h, P[f"{n}.prelu.alpha"])))
        return ops.prelu(ops.conv3d(x, emu.w(P["head.kernel"]), P["head.bias"]), P["head.prelu.alpha"])


# ------------------------------------------------------------------ decoder D5 (vqgan_attn_cp)
class AttnCpDecoder:
    """vqgan_attn_cp.Decoder: Conv1(D->c_top) GN(min(D,32)) SiLU; [ConvT(k4,s2,c_i) 2xResUnit]x(L-1); Conv3(->out).
    ResUnit: GN(min(C,32),eps 1e-6) SiLU Conv3 GN SiLU Conv3 + x."""

    def __init__(self, in_channels, out_channels, num_channels):
        self.cin, self.cout = in_channels, out_channels
        self.ch = list(reversed(num_channels))

    def spec(self):
        c0 = self.ch[0]
        sp = [("stem.kernel", (1, 1, 1, self.cin, c0), "glorot"), ("stem.bias", (c0,), "zeros"),
              ("stem.norm.gamma", (c0,), "ones"), ("stem.norm.beta", (c0,), "zeros")]
        for i in range(1, len(self.ch)):
            c = self.ch[i]
            sp += [(f"level.{i}.up.kernel", (4, 4, 4, c, self.ch[i - 1]), "glorot"), (f"level.{i}.up.bias", (c,), "zeros")]
            for j in range(2):
                n = f"level.{i}.res.{j}"
                sp += [(f"{n}.norm1.gamma", (c,), "ones"), (f"{n}.norm1.beta", (c,), "zeros"),
                       (f"{n}.conv1.kernel", (3, 3, 3, c, c), "glorot"), (f"{n}.conv1.bias", (c,), "zeros"),
                       (f"{n}.norm2.gamma", (c,), "ones"), (f"{n}.norm2.beta", (c,), "zeros"),
                       (f"{n}.conv2.kernel", (3, 3, 3, c, c), "glorot"), (f"{n}.conv2.bias", (c,), "zeros")]
        sp += [("head.kernel", (3, 3, 3, self.ch[-1], self.cout), "glorot"), ("head.bias", (self.cout,), "zeros")]
        return sp

    def stem_groups(self):
        # GroupNormalization(groups=min(self.in_channels, 32)) applied to the c_top-channel tensor
        return min(self.cin, 32)

    def forward(self, P, z, emu: Emu = EXACT):
        x = emu.a(ops.conv3d(emu.a(z), emu.w(P["stem.kernel"]), P["stem.bias"]))
        x = emu.a(ops.swish(ops.groupnorm(x, P["stem.norm.gamma"], P["stem.norm.beta"], self.stem_groups(), 1e-6)))
        for i in range(1, len(self.ch)):
            c = self.ch[i]
            g = min(c, 32)
            x = emu.a(ops.conv3d_transpose(x, emu.w(P[f"level.{i}.up.kernel"]), P[f"level.{i}.up.bias"]))
            for j in range(2):
                n = f"level.{i}.res.{j}"
                h = emu.a(ops.swish(ops.groupnorm(x, P[f"{n}.norm1.gamma"], P[f"{n}.norm1.beta"], g, 1e-6)))
                h = emu.a(ops.conv3d(h, emu.w(P[f"{n}.conv1.kernel"]), P[f"{n}.conv1.bias"]))
                h = emu.a(ops.swish(ops.groupnorm(h, P[f"{n}.norm2.gamma"], P[f"{n}.norm2.beta"], g, 1e-6)))
                x = emu.a(ops.conv3d(h, emu.w(P[f"{n}.conv2.kernel"]), P[f"{n}.conv2.bias"]) + x)
        return ops.conv3d(x, emu.w(P["head.kernel"]), P["head.bias"])


# ------------------------------------------------------------------ parameter counting (known-answer pin)
def monai_vqvae_param_count(in_channels, out_channels, num_channels, num_res_layers, num_res_channels,
                            num_embeddings, embedding_dim, img_size=128):
    """(trainable, bn_moving_stats) of vqvae3d_monai.VQVAE (encoder :237-306, decoder :309-391, quantizer
    :112-131) with down/upsample parameters (2,4,1,'same'); per-voxel PReLU alphas; biases everywhere."""
    tr, nt = 0, 0
    s, c_prev = img_size, in_channels
    for i, c in enumerate(num_channels):  # encoder
        tr += 4 ** 3 * c_prev * c + c
        s //= 2
        for _ in range(num_res_layers):
            rc = num_res_channels[i]
            tr += 27 * c * rc + rc + 27 * rc * c + c + 2 * c + s ** 3 * c
            nt += 2 * c
        c_prev = c
    tr += 27 * c_prev * embedding_dim + embedding_dim + s ** 3 * embedding_dim
    dec = MonaiDecoder(embedding_dim, out_channels, num_channels, num_res_layers, num_res_channels, s)
    for name, shp, _ in dec.spec():
        n = int(np.prod(shp))
        if name.endswith(".mean") or name.endswith(".var"):
            nt += n
        else:
            tr += n
    tr += embedding_dim * num_embeddings
    return tr, nt


# ------------------------------------------------------------------ decoders D2 / D3 / D4 (vqgan, vqgan_gnorm, vqgan_stride)
class VqganFamilyDecoder:
    """The three VQ-GAN decoder variants that share one block list (act_fn='prelu', the scripts' default):

      'vqgan'  (D2, networks/vqgan.py:378-475, res unit :256-284)   stem Conv3 -> BN -> PReLU;  ConvT(k4,s2) -> BN [-> PReLU]
      'gnorm'  (D3, networks/vqgan_gnorm.py:382-484, :256-286)       BN -> GroupNorm: GN(8, eps 1e-6); res unit GN(groups=1,
               default eps 1e-3) when its width is 2; after ConvT GN(int(out/2), default eps 1e-3) when out < 32
      'stride' (D4, networks/vqgan_stride.py:376-480, :256-286)      stem Conv3 -> PReLU (no norm); res unit with BN;
               Conv3D(k=4, s=1, 'same' = pad (1,2)) -> UpSampling3D(2) -> [GN(int(out/2), eps 1e-6) if out < 32] [-> PReLU]
    Res unit everywhere: relu(x + PReLU(Norm(Conv3(relu(Conv3(x)))))).  PReLU has one alpha per (d,h,w,c) element.
    The level's activation is skipped after the last level; `output_act` appends a ReLU."""

    def __init__(self, variant, in_channels, out_channels, num_channels, num_res_layers, num_res_channels, in_size,
                 output_act=None):
        assert variant in ("vqgan", "gnorm", "stride")
        self.variant, self.cin, self.cout = variant, in_channels, out_channels
        self.ch, self.rch = list(reversed(num_channels)), list(reversed(num_res_channels))
        self.R, self.s0, self.output_act = num_res_layers, in_size, output_act

    # (kind, groups, eps) of the normalisation at each site; kind None = no norm
    def stem_norm(self):
        return {"vqgan": ("bn", 0, 1e-3), "gnorm": ("gn", 8, 1e-6), "stride": (None, 0, 0.0)}[self.variant]

    def res_norm(self, c):
        if self.variant == "gnorm":
            return ("gn", 1, 1e-3) if c == 2 else ("gn", 8, 1e-6)
        return ("bn", 0, 1e-3)

    def up_norm(self, out):
        if self.variant == "vqgan":
            return ("bn", 0, 1e-3)
        if self.variant == "gnorm":
            return ("gn", int(out / 2), 1e-3) if out < 32 else ("gn", 8, 1e-6)
        return ("gn", int(out / 2), 1e-6) if out < 32 else (None, 0, 0.0)

    @staticmethod
    def _norm_spec(name, kind, c):
        if kind is None:
            return []
        sp = [(f"{name}.gamma", (c,), "ones"), (f"{name}.beta", (c,), "zeros")]
        if kind == "bn":
            sp += [(f"{name}.mean", (c,), "zeros"), (f"{name}.var", (c,), "ones")]
        return sp

    def spec(self):
        s, c = self.s0, self.ch[0]
        sp = [("stem.kernel", (3, 3, 3, self.cin, c), "glorot"), ("stem.bias", (c,), "zeros")]
        sp += self._norm_spec("stem.norm", self.stem_norm()[0], c) + [("stem.prelu.alpha", (s, s, s, c), "zeros")]
        for i, c in enumerate(self.ch):
            for j in range(self.R):
                n, rc = f"level.{i}.res.{j}", self.rch[i]
                sp += [(f"{n}.conv1.kernel", (3, 3, 3, c, rc), "glorot"), (f"{n}.conv1.bias", (rc,), "zeros"),
                       (f"{n}.conv2.kernel", (3, 3, 3, rc, c), "glorot"), (f"{n}.conv2.bias", (c,), "zeros")]
                sp += self._norm_spec(f"{n}.norm", self.res_norm(c)[0], c) + [(f"{n}.prelu.alpha", (s, s, s, c), "zeros")]
            last = i == len(self.ch) - 1
            out = self.cout if last else self.ch[i + 1]
            if self.variant == "stride":
                sp += [(f"level.{i}.up.kernel", (4, 4, 4, c, out), "glorot")]
            else:
                sp += [(f"level.{i}.up.kernel", (4, 4, 4, out, c), "glorot")]
            sp += [(f"level.{i}.up.bias", (out,), "zeros")]
            s *= 2
            sp += self._norm_spec(f"level.{i}.up.norm", self.up_norm(out)[0], out)
            if not last:
                sp += [(f"level.{i}.up.prelu.alpha", (s, s, s, out), "zeros")]
        return sp

    @staticmethod
    def _norm(P, name, x, kind, groups, eps):
        if kind is None:
            return x
        if kind == "bn":
            return ops.batchnorm_infer(x, P[f"{name}.gamma"], P[f"{name}.beta"], P[f"{name}.mean"], P[f"{name}.var"], eps)
        return ops.groupnorm(x, P[f"{name}.gamma"], P[f"{name}.beta"], groups, eps)

    def forward(self, P, z, emu: Emu = EXACT):
        x = ops.conv3d(emu.a(z), emu.w(P["stem.kernel"]), P["stem.bias"])
        x = emu.a(ops.prelu(self._norm(P, "stem.norm", x, *self.stem_norm()), P["stem.prelu.alpha"]))
        for i, c in enumerate(self.ch):
            for j in range(self.R):
                n = f"level.{i}.res.{j}"
                h = emu.a(torch.relu(ops.conv3d(x, emu.w(P[f"{n}.conv1.kernel"]), P[f"{n}.conv1.bias"])))
                h = ops.conv3d(h, emu.w(P[f"{n}.conv2.kernel"]), P[f"{n}.conv2.bias"])
                h = self._norm(P, f"{n}.norm", h, *self.res_norm(c))
                x = emu.a(torch.relu(x + ops.prelu(h, P[f"{n}.prelu.alpha"])))
            last = i == len(self.ch) - 1
            out = self.cout if last else self.ch[i + 1]
            if self.variant == "stride":
                x = ops.upsample_nearest2(ops.conv3d(x, emu.w(P[f"level.{i}.up.kernel"]), P[f"level.{i}.up.bias"]))
            else:
                x = ops.conv3d_transpose(x, emu.w(P[f"level.{i}.up.kernel"]), P[f"level.{i}.up.bias"])
            x = self._norm(P, f"level.{i}.up.norm", x, *self.up_norm(out))
            if not last:
                x = emu.a(ops.prelu(x, P[f"level.{i}.up.prelu.alpha"]))
        if self.output_act:
            x = torch.relu(x)
        return x
