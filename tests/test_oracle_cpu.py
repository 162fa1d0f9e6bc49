"""CPU tests of the oracle (test infrastructure): known-answer pins and the committed golden vectors.

The reference ships no tests or golden vectors for this path and TensorFlow cannot run here, so the pins are
  * published known answers of the algorithms the oracle restates (Random123 Philox4x32-10 vectors);
  * the one machine-checkable number the reference does ship: the Keras parameter count of the monai VQ-VAE
    (experiments/vqvae3d-scaled-monai-B8-AUG-all-T-KR.output:23-25);
  * closed-form identities of `Betas` (dm3d.py:194-214) and of DiffusionModel.sample (dm3d.py:477-508);
  * Keras op semantics checked against independent direct-loop evaluations;
  * the seeded golden vectors under tests/golden/ (tools/make_golden.py) -- regression pins of the oracle itself.
"""
import os

import numpy as np
import pytest
import torch

from oracle import first_stage as OF, init as OI, ops as O, philox, sampler as OS
from oracle.schedule import Betas
from oracle.unet import UNet

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
torch.set_num_threads(min(8, os.cpu_count() or 1))


def close(a, b, rel=2e-5):
    """fp32 summation order depends on the host thread count: compare by rel-L2 and a scaled max-abs."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item() <= rel and (a - b).abs().max().item() <= 50 * rel * b.abs().max().item()


# ----------------------------------------------------------------------------------------- Philox
def test_philox_random123_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
        assert tuple(int(g) for g in got) == want


def test_philox_normal_golden_and_moments():
    z = philox.normal(1234, 17, [5, 6, 7], 4 * 4 * 4 * 8)
    g = np.load(os.path.join(GOLD, "philox.npz"))["z"]
    assert np.array_equal(z, g)
    big = philox.normal(99, 3, np.arange(4), 1 << 16)
    assert abs(big.mean()) < 1e-2 and abs(big.std() - 1) < 1e-2
    assert np.isfinite(big).all()
    # counter layout: sample id and step select independent streams; prefix property over n_elem
    assert np.array_equal(philox.normal(99, 3, [2], 100), big[2:3, :100])
    assert not np.array_equal(philox.normal(99, 4, [2], 100), big[2:3, :100])


# ----------------------------------------------------------------------------------------- schedule / update
@pytest.mark.parametrize("T", [50, 300, 1000])
def test_betas_identities(T):
    b = Betas(T)
    assert b.beta.dtype == np.float32 and b.beta.shape == (T,)
    assert np.isclose(b.beta[0], 1e-4) and np.isclose(b.beta[-1], 2e-2)
    ab64 = np.cumprod(1 - np.linspace(1e-4, 2e-2, T))
    assert np.allclose(b.alpha_bar, ab64, rtol=1e-6)
    assert b.alpha_bar_prev[0] == 1.0 and np.array_equal(b.alpha_bar_prev[1:], b.alpha_bar[:-1])
    assert np.allclose(b.sqrt_alpha_bar ** 2, b.alpha_bar, rtol=1e-6)
    assert np.allclose(b.sqrt_one_minus_alpha_bar ** 2 + b.alpha_bar, 1.0, rtol=1e-6)
    assert np.all(np.diff(b.alpha_bar) < 0)


def test_sample_closed_form_and_last_step():
    b = Betas(100)
    x = OI.normal((2, 4, 4, 4, 3), 1)
    e = OI.normal((2, 4, 4, 4, 3), 2)
    # posterior mean == (x_t - beta/sqrt(1-ab) eps)/sqrt(alpha) (Ho et al. eq. 11), to fp32 rounding
    for t in (99, 50, 1):
        mean, var = OS.sample(b, x, e, t)
        alt = (x - b.beta[t] / b.sqrt_one_minus_alpha_bar[t] * e) / b.sqrt_alpha[t]
        assert torch.allclose(mean, alt, rtol=2e-3, atol=2e-4)
        assert var > 0
    # t == 0: variance is exactly 0 -> sigma = sqrt(1e-20), noise ignored, mean clipped
    m0, v0 = OS.sample(b, x, e, 0)
    assert float(v0) == 0.0
    out = OS.ddpm_step(b, x, e, 0, noise=torch.ones_like(x))
    assert torch.equal(out, m0.clamp(-1, 1))
    out5 = OS.ddpm_step(b, 5 * x, e, 50, noise=None)
    assert out5.abs().max() <= 1.0


def test_ddim_extension_is_deterministic_and_reaches_x0():
    b = Betas(100)
    x = OI.normal((1, 2, 2, 2, 3), 1)
    e = OI.normal((1, 2, 2, 2, 3), 2)
    x0 = ((x - b.sqrt_one_minus_alpha_bar[10] * e) / b.sqrt_alpha_bar[10]).clamp(-1, 1)
    assert torch.allclose(OS.ddim_step(b, x, e, 10, -1), x0)
    assert torch.equal(OS.ddim_step(b, x, e, 10, 5), OS.ddim_step(b, x, e, 10, 5))


# ----------------------------------------------------------------------------------------- parameter-count pin
def test_monai_vqvae_param_count_matches_reference_log():
    """experiments/vqvae3d-scaled-monai-B8-AUG-all-T-KR.output:23-25: 75,593,473 trainable; 2,694 non-trainable
    = 2,688 BN moving statistics + 6 metric variables."""
    tr, nt = OF.monai_vqvae_param_count(1, 1, (32, 64, 128), 3, (32, 64, 128), 256, 64, img_size=128)
    assert tr == 75_593_473
    assert nt == 2_688
    # the same number from the encoder / decoder weight tables the oracle and the product are built from
    enc = OF.MonaiEncoder(1, 64, (32, 64, 128), 3, (32, 64, 128), 128)
    dec = OF.MonaiDecoder(64, 1, (32, 64, 128), 3, (32, 64, 128), 16)
    stats = lambda n: n.endswith(".mean") or n.endswith(".var")  # noqa: E731
    tr2 = sum(int(np.prod(s)) for sp in (enc.spec(), dec.spec()) for n, s, _ in sp if not stats(n)) + 64 * 256
    nt2 = sum(int(np.prod(s)) for sp in (enc.spec(), dec.spec()) for n, s, _ in sp if stats(n))
    assert (tr2, nt2) == (75_593_473, 2_688)


def test_monai_encoder_shapes_and_product_spec_agree():
    import b200dm
    oe = OF.MonaiEncoder(1, 8, (32, 64), 1, (32, 64), 16)
    P = OI.make_params(oe.spec(), 6, "stress")
    z = oe.forward(P, OI.normal((1, 16, 16, 16, 1), 12))
    assert tuple(z.shape) == (1, 4, 4, 4, 8)
    pe = b200dm.MonaiEncoder(1, 8, (32, 64), 1, (32, 64), 16)
    assert [(n, tuple(s)) for n, s, _ in pe.spec] == [(n, tuple(s)) for n, s, _ in oe.spec()]


def test_unet_param_counts_are_stable():
    u = UNet(16, 8, [64, 128, 256], [False, False, True, True], first_conv_channels=64)
    n = sum(int(np.prod(s)) for _, s, _ in u.spec())
    uc = UNet(32, 256, [64, 128, 256], [False, False, True, True], first_conv_channels=32, conditional=True)
    nc = sum(int(np.prod(s)) for _, s, _ in uc.spec())
    # ~42 M unconditional; 47-146 M conditional depending on ContextMLP size (6 x 16.8 M here)
    assert 40e6 < n < 45e6, n
    assert 140e6 < nc < 150e6, nc


# ----------------------------------------------------------------------------------------- Keras op semantics
def test_same_padding_rules():
    assert O.same_pads(32, 3, 1) == (1, 1)
    assert O.same_pads(32, 3, 2) == (0, 1)      # asymmetric: NOT torch padding=1
    assert O.same_pads(32, 4, 1) == (1, 2)
    assert O.same_pads(7, 3, 2) == (1, 1)
    assert O.same_pads(32, 1, 1) == (0, 0)


def _conv_direct(x, k, stride):
    """Independent direct-loop evaluation of a channels-last cross-correlation with TF 'same' padding."""
    N, D, H, W, Ci = x.shape
    kk, Co = k.shape[0], k.shape[-1]
    pb = [O.same_pads(s, kk, stride)[0] for s in (D, H, W)]
    od, oh, ow = [-(-s // stride) for s in (D, H, W)]
    y = np.zeros((N, od, oh, ow, Co), np.float64)
    xn, kn = x.numpy().astype(np.float64), k.numpy().astype(np.float64)
    for a in range(od):
        for b_ in range(oh):
            for c in range(ow):
                for i in range(kk):
                    for j in range(kk):
                        for l in range(kk):
                            d, h, w = a * stride + i - pb[0], b_ * stride + j - pb[1], c * stride + l - pb[2]
                            if 0 <= d < D and 0 <= h < H and 0 <= w < W:
                                y[:, a, b_, c, :] += xn[:, d, h, w, :] @ kn[i, j, l]
    return y


@pytest.mark.parametrize("k,stride", [(3, 1), (3, 2), (1, 1), (4, 1)])
def test_conv3d_against_direct_loops(k, stride):
    x = OI.normal((1, 4, 6, 4, 3), 1)
    w = OI.normal((k, k, k, 3, 5), 2)
    y = O.conv3d(x, w, None, stride=stride).numpy()
    assert np.allclose(y, _conv_direct(x, w, stride), atol=1e-4)


def test_conv3d_transpose_against_scatter_definition():
    """Conv3DTranspose(k=4,s=2,'same'): out[2i + t - 1] += in[i] * w[t] (kernel (kd,kh,kw,Cout,Cin)), out = 2*in."""
    x = OI.normal((1, 3, 3, 3, 2), 1)
    w = OI.normal((4, 4, 4, 5, 2), 2)
    y = O.conv3d_transpose(x, w).numpy()
    want = np.zeros((1, 6, 6, 6, 5))
    xn, wn = x.numpy().astype(np.float64), w.numpy().astype(np.float64)
    for i in range(3):
        for j in range(3):
            for l in range(3):
                for a in range(4):
                    for b_ in range(4):
                        for c in range(4):
                            d, h, ww = 2 * i + a - 1, 2 * j + b_ - 1, 2 * l + c - 1
                            if 0 <= d < 6 and 0 <= h < 6 and 0 <= ww < 6:
                                want[0, d, h, ww] += wn[a, b_, c] @ xn[0, i, j, l]
    assert y.shape == want.shape and np.allclose(y, want, atol=1e-4)


def test_norm_semantics():
    x = OI.normal((2, 3, 3, 3, 8), 1)
    g, b = torch.rand(8) + 0.5, torch.randn(8)
    m, v = torch.randn(8) * 0.1, torch.rand(8) + 0.5
    bn = O.batchnorm_infer(x, g, b, m, v)
    assert torch.allclose(bn, g * (x - m) / torch.sqrt(v + 1e-3) + b, atol=1e-5)
    gn = O.groupnorm(x, g, b, 4, 1e-6)
    ref = torch.nn.functional.group_norm(x.permute(0, 4, 1, 2, 3), 4, g, b, 1e-6).permute(0, 2, 3, 4, 1)
    assert torch.allclose(gn, ref, atol=1e-5)
    ln = O.layernorm(x, g, b)
    assert torch.allclose(ln, torch.nn.functional.layer_norm(x, (8,), g, b, 1e-3), atol=1e-5)
    a = torch.rand(3, 3, 3, 8)
    assert torch.equal(O.prelu(x, a), torch.where(x > 0, x, a * x))
    u = O.upsample_nearest2(x)
    assert u.shape == (2, 6, 6, 6, 8) and torch.equal(u[:, 3, 4, 5], x[:, 1, 2, 2])


def test_time_embedding():
    e = O.time_embedding(torch.tensor([0, 7]), 256)
    assert e.shape == (2, 256)
    assert torch.equal(e[0, :128], torch.zeros(128)) and torch.equal(e[0, 128:], torch.ones(128))
    assert np.isclose(float(e[1, 0]), np.sin(7.0), atol=1e-6) and np.isclose(float(e[1, 127]), np.sin(7.0 / 10000), atol=1e-7)


# ----------------------------------------------------------------------------------------- VQ
def test_vq_first_min_on_ties_and_layouts():
    cb = OI.codebook(16, 8, "KD", seed=3)
    cb[9] = cb[4]                                   # duplicate code: the lower index must win
    x = cb[[4, 9, 0, 15]] + 1e-4
    idx = OF.get_code_indices(x, cb, "KD")
    assert idx.tolist() == [4, 4, 0, 15] and idx.dtype == torch.int64
    assert torch.equal(OF.get_code_indices(x, cb.t().contiguous(), "DK"), idx)
    q, idx2, perp, counts = OF.quantize(x.reshape(1, 1, 2, 2, 8), cb, "KD")
    assert torch.equal(q.reshape(-1, 8), cb[idx]) and counts.sum() == 4 and counts[4] == 2
    p = counts.double() / 4
    assert np.isclose(float(perp), float(torch.exp(-(p * torch.log(p + 1e-10)).sum())))
    ie, margin = OF.get_code_indices_exact(x, cb, "KD")
    assert torch.equal(ie, idx) and margin.min() >= 0


def test_vq_empty_input():
    cb = OI.codebook(16, 8, "KD", seed=3)
    idx = OF.get_code_indices(torch.zeros(0, 8), cb, "KD")
    assert idx.shape == (0,)


# ----------------------------------------------------------------------------------------- golden vectors
def test_golden_unet_uncond():
    g = np.load(os.path.join(GOLD, "unet_uncond_8.npz"))
    u = UNet(8, 8, [64, 128, 256], [False, False, True, True], first_conv_channels=64)
    P = OI.make_params(u.spec(), 0, "stress")
    x = torch.from_numpy(g["x"])
    assert torch.equal(x, OI.normal((2, 8, 8, 8, 8), 1))
    eps = u.forward(P, x, torch.tensor([int(g["t"])] * 2))
    assert close(eps, g["eps"])
    # independent high-precision evaluation: fp64 arithmetic on the same weights
    P64 = {k: v.double() for k, v in P.items()}
    eps64 = u.forward(P64, x.double(), torch.tensor([int(g["t"])] * 2))
    r = ((eps.double() - eps64).norm() / eps64.norm()).item()
    assert r < 1e-5, r


def test_golden_unet_cond():
    g = np.load(os.path.join(GOLD, "unet_cond_8.npz"))
    u = UNet(8, 16, [64, 128, 256], [False, False, True, True], first_conv_channels=32, conditional=True)
    P = OI.make_params(u.spec(), 0, "stress")
    eps = u.forward(P, torch.from_numpy(g["x"]), torch.tensor([int(g["t"])] * 2), ctx=torch.from_numpy(g["ctx"]))
    assert close(eps, g["eps"])
    # the context must matter (cross-attention is live) and samples must be independent of their batch-mates
    eps_sw = u.forward(P, torch.from_numpy(g["x"]), torch.tensor([int(g["t"])] * 2), ctx=torch.tensor([1, 0]))
    assert not torch.allclose(eps_sw, eps, atol=1e-4)
    one = u.forward(P, torch.from_numpy(g["x"][1:]), torch.tensor([int(g["t"])]), ctx=torch.tensor([1]))
    assert close(one, eps[1:])


def test_golden_chain():
    g = np.load(os.path.join(GOLD, "chain_T12.npz"))
    u = UNet(8, 8, [64, 128, 256], [False, False, True, True], first_conv_channels=64)
    P = OI.make_params(u.spec(), 0, "stress")
    T, shape = 12, (2, 8, 8, 8, 8)
    noises = {i: OI.normal(shape, 100 + i) for i in range(1, T)}
    lat = OS.generate(lambda z, i: u.forward(P, z, torch.full((2,), i)), Betas(T), OI.normal(shape, 1234), noises=noises)
    assert close(lat, g["latents"])
    assert lat.abs().max() <= 1.0 + 1e-6            # the last step returns the clipped mean


def test_golden_vq_and_decoders():
    g = np.load(os.path.join(GOLD, "vq_decode.npz"))
    cb = OI.codebook(256, 64, "KD", seed=3)
    z = OI.normal((2, 4, 4, 4, 64), 11, 0.05)
    q, idx, perp, counts = OF.quantize(z, cb, "KD")
    assert np.array_equal(idx.numpy(), g["idx"])                       # bit-exact
    assert np.isclose(float(perp), float(g["perplexity"]), rtol=1e-6)
    dec = OF.AttnCpDecoder(64, 1, (32, 64, 128))
    vol = dec.forward(OI.make_params(dec.spec(), 5, "stress"), q)
    assert vol.shape == (2, 16, 16, 16, 1)
    assert close(vol, g["vol_attn_cp"])
    md = OF.MonaiDecoder(64, 1, (32, 64), 2, (32, 64), 4)
    volm = md.forward(OI.make_params(md.spec(), 5, "stress"), z)
    assert volm.shape == (2, 16, 16, 16, 1)
    assert close(volm, g["vol_monai"])


def test_bf16_emulation_is_close_to_exact():
    """Emu(True) rounds at the CUDA path's storage points; the induced error is the stated bf16 budget."""
    u = UNet(8, 8, [64, 128, 256], [False, False, True, True], first_conv_channels=64)
    P = OI.make_params(u.spec(), 0, "stress")
    x = OI.normal((1, 8, 8, 8, 8), 1)
    t = torch.tensor([5])
    a, b = u.forward(P, x, t), u.forward(P, x, t, emu=O.Emu(True))
    r = ((a - b).norm() / a.norm()).item()
    assert 0 < r < 2.5e-2, r


@pytest.mark.parametrize("variant", ["vqgan", "gnorm", "stride"])
def test_vqgan_family_oracle_structure(variant):
    """D2-D4 oracle: output doubles per level, keras-init (alpha = 0, gamma = 1) makes PReLU a ReLU, and with zero conv
    weights the output is the norm/bias chain of a zero tensor (finite, sample-independent)."""
    d = OF.VqganFamilyDecoder(variant, 16, 2, (32, 64), 1, (32, 64), 4)
    P = OI.make_params(d.spec(), 5, "stress")
    z = OI.normal((2, 4, 4, 4, 16), 1, 0.5)
    y = d.forward(P, z)
    assert y.shape == (2, 16, 16, 16, 2) and torch.isfinite(y).all()
    one = d.forward(P, z[1:])
    assert close(one, y[1:], rel=1e-5)                      # samples are independent (GN is per sample, BN is inference)
    Pk = OI.make_params(d.spec(), 5, "keras")
    assert all(float(v.abs().max()) == 0 for k, v in Pk.items() if k.endswith("alpha"))
