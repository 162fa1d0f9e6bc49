"""3D latent U-Net (oracle; test infrastructure only).

Restates ``build_model`` and its blocks:
  unconditional  networks/dm3d.py:294-376   (ResidualBlock :217-252, DownSample :255-266,
                 UpSample :269-277, TimeMLP :280-288, TimeEmbedding :177-191, AttentionBlock :18-63)
  conditional    networks/conditional_dm3d.py:324-415 (CrossAttentionBlock :112-195,
                 ContextMLP :310-318, Embedding :358)
Semantics kept literally: BN in inference mode with eps 1e-3; AttentionBlock
returns BN(x)+proj (the normalised input is the residual); CrossAttentionBlock's three
branches all read ``h``; q/k/v Dense weights are shared by self- and cross-attention;
attention scale = units^-1/2; ``widths[i] != widths[-1]`` decides down-sampling by VALUE;
up-path concat order is [x, skip].

Parameters live in a flat dict {canonical_name: tensor in Keras layout}; ``spec()`` lists
(name, shape, initialiser) in the reference's layer-construction order.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import Emu, EXACT


class UNet:
    def __init__(self, img_size, img_channels, widths, has_attention, num_res_blocks=2,
                 first_conv_channels=64, conditional=False, context_dim=1):
        self.S, self.C_lat, self.widths = img_size, img_channels, list(widths)
        self.has_attention = list(has_attention)
        self.R, self.F, self.cond, self.context_dim = num_res_blocks, first_conv_channels, conditional, context_dim
        self._spec = []
        self._build()

    # ---- structure walk (mirrors build_model line by line) -------------------------
    def _p(self, name, shape, init):
        self._spec.append((name, tuple(shape), init))

    def _res(self, name, cin, w):
        if cin != w:
            self._p(f"{name}.shortcut.kernel", (1, 1, 1, cin, w), "vs1"); self._p(f"{name}.shortcut.bias", (w,), "zeros")
        self._p(f"{name}.temb.kernel", (4 * self.F, w), "vs1"); self._p(f"{name}.temb.bias", (w,), "zeros")
        self._bn(f"{name}.norm1", cin)
        self._p(f"{name}.conv1.kernel", (3, 3, 3, cin, w), "vs1"); self._p(f"{name}.conv1.bias", (w,), "zeros")
        self._bn(f"{name}.norm2", w)
        self._p(f"{name}.conv2.kernel", (3, 3, 3, w, w), "vs0"); self._p(f"{name}.conv2.bias", (w,), "zeros")

    def _bn(self, name, c):
        self._p(f"{name}.gamma", (c,), "ones"); self._p(f"{name}.beta", (c,), "zeros")
        self._p(f"{name}.mean", (c,), "zeros"); self._p(f"{name}.var", (c,), "ones")

    def _dense(self, name, cin, cout, init="vs1"):
        self._p(f"{name}.kernel", (cin, cout), init); self._p(f"{name}.bias", (cout,), "zeros")

    def _attn(self, name, c, s):
        if not self.cond:  # AttentionBlock (dm3d.py:26-37); self.depth Dense is never called -> no weights
            self._bn(f"{name}.norm", c)
            for n in ("query", "key", "value"):
                self._dense(f"{name}.{n}", c, c)
            self._dense(f"{name}.proj", c, c, "vs0")
        else:  # ContextMLP (conditional_dm3d.py:310-318) then CrossAttentionBlock (:120-138)
            self._dense(f"{name}.ctxmlp", 4 * self.F, s * s * s * c, "glorot")
            self._bn(f"{name}.norm", c)
            self._p(f"{name}.proj_in.kernel", (1, 1, 1, c, c), "glorot"); self._p(f"{name}.proj_in.bias", (c,), "zeros")
            for n in ("norm1", "norm2", "norm3"):
                self._p(f"{name}.{n}.gamma", (c,), "ones"); self._p(f"{name}.{n}.beta", (c,), "zeros")
            for n in ("query", "key", "value"):
                self._dense(f"{name}.{n}", c, c, "glorot")
            self._dense(f"{name}.mlp0", c, 4 * c, "glorot")
            self._dense(f"{name}.mlp1", 4 * c, c, "glorot")
            self._p(f"{name}.proj_out.kernel", (1, 1, 1, c, c), "glorot"); self._p(f"{name}.proj_out.bias", (c,), "zeros")

    def _build(self):
        F, W = self.F, self.widths
        self.prog = []  # (kind, name, args...)
        self._p("in.kernel", (3, 3, 3, self.C_lat, F), "vs1"); self._p("in.bias", (F,), "zeros")
        self._dense("time.dense0", 4 * F, 4 * F); self._dense("time.dense1", 4 * F, 4 * F)
        if self.cond:
            self._p("ctx.embedding", (self.context_dim + 1, 4 * F), "embed")
        c, s = F, self.S
        skips = [c]
        for i, w in enumerate(W):
            for j in range(self.R):
                self._res(f"down.{i}.res.{j}", c, w); self.prog.append(("res", f"down.{i}.res.{j}", c, w)); c = w
                if self.has_attention[i]:
                    self._attn(f"down.{i}.attn.{j}", c, s); self.prog.append(("attn", f"down.{i}.attn.{j}", c, s))
                skips.append(c); self.prog.append(("push",))
            if w != W[-1]:
                self._p(f"down.{i}.downsample.kernel", (3, 3, 3, c, w), "vs1"); self._p(f"down.{i}.downsample.bias", (w,), "zeros")
                self.prog.append(("down", f"down.{i}.downsample")); c = w; s //= 2
                skips.append(c); self.prog.append(("push",))
        self._res("mid.res.0", c, W[-1]); self.prog.append(("res", "mid.res.0", c, W[-1])); c = W[-1]
        self._attn("mid.attn", c, s); self.prog.append(("attn", "mid.attn", c, s))
        self._res("mid.res.1", c, c); self.prog.append(("res", "mid.res.1", c, c))
        for i in reversed(range(len(W))):
            w = W[i]
            for j in range(self.R + 1):
                cs = skips.pop()
                self.prog.append(("cat",))
                self._res(f"up.{i}.res.{j}", c + cs, w); self.prog.append(("res", f"up.{i}.res.{j}", c + cs, w)); c = w
                if self.has_attention[i]:
                    self._attn(f"up.{i}.attn.{j}", c, s); self.prog.append(("attn", f"up.{i}.attn.{j}", c, s))
            if i != 0:
                self._p(f"up.{i}.upsample.kernel", (3, 3, 3, c, w), "vs1"); self._p(f"up.{i}.upsample.bias", (w,), "zeros")
                self.prog.append(("up", f"up.{i}.upsample")); s *= 2
        self._bn("out.norm", c)
        self._p("out.conv.kernel", (3, 3, 3, c, self.C_lat), "vs0"); self._p("out.conv.bias", (self.C_lat,), "zeros")

    def spec(self):
        return list(self._spec)

    # ---- forward -------------------------------------------------------------------
    def temb(self, P, t):
        """TimeEmbedding -> TimeMLP (dm3d.py:325-326): (B,) int -> (B,4F) fp32."""
        e = ops.time_embedding(t, 4 * self.F).to(P["time.dense0.kernel"].dtype)  # float64 when the oracle runs in its high-precision mode
        e = ops.swish(ops.dense(e, P["time.dense0.kernel"], P["time.dense0.bias"]))
        return ops.dense(e, P["time.dense1.kernel"], P["time.dense1.bias"])

    def _bn_apply(self, P, name, x):
        return ops.batchnorm_infer(x, P[f"{name}.gamma"], P[f"{name}.beta"], P[f"{name}.mean"], P[f"{name}.var"])

    def _resblock(self, P, name, x, temb, cin, w, emu: Emu):
        if cin == w:
            res = x
        else:
            res = emu.a(ops.conv3d(x, emu.w(P[f"{name}.shortcut.kernel"]), P[f"{name}.shortcut.bias"]), f"{name}.shortcut")
        e = ops.dense(ops.swish(temb), P[f"{name}.temb.kernel"], P[f"{name}.temb.bias"])[:, None, None, None, :]
        h = emu.a(ops.swish(self._bn_apply(P, f"{name}.norm1", x)), f"{name}.norm1")
        h = emu.a(ops.conv3d(h, emu.w(P[f"{name}.conv1.kernel"]), P[f"{name}.conv1.bias"]) + e, f"{name}.conv1")
        h = emu.a(ops.swish(self._bn_apply(P, f"{name}.norm2", h)), f"{name}.norm2")
        return emu.a(ops.conv3d(h, emu.w(P[f"{name}.conv2.kernel"]), P[f"{name}.conv2.bias"]) + res, f"{name}.conv2")

    def _d(self, P, name, x, emu, act=None):
        y = ops.dense(x, emu.w(P[f"{name}.kernel"]), P[f"{name}.bias"])
        if act == "relu":
            y = torch.relu(y)
        return y

    def _attention(self, P, name, x, c, emu: Emu):
        """AttentionBlock.call (dm3d.py:39-63)."""
        B, s = x.shape[0], x.shape[1]
        n = emu.a(self._bn_apply(P, f"{name}.norm", x), f"{name}.norm")
        f = n.reshape(B, s ** 3, c)
        q, k, v = (emu.a(self._d(P, f"{name}.{m}", f, emu)) for m in ("query", "key", "value"))
        o = emu.a(ops.attention_core(q, k, v, float(c) ** -0.5, emu))
        return emu.a(n + self._d(P, f"{name}.proj", o, emu).reshape(x.shape), f"{name}.proj")

    def context_kv(self, P, name, cemb, c, s, emu: Emu):
        """ContextMLP (conditional_dm3d.py:310-318) + key(ctx), value(ctx) (:168-169): step-invariant."""
        ctx = emu.a(ops.swish(ops.dense(cemb, P[f"{name}.ctxmlp.kernel"], P[f"{name}.ctxmlp.bias"])))
        ctx = ctx.reshape(-1, s ** 3, c)
        return emu.a(self._d(P, f"{name}.key", ctx, emu)), emu.a(self._d(P, f"{name}.value", ctx, emu))

    def _xattention(self, P, name, x, cemb, c, s, emu: Emu):
        """CrossAttentionBlock.call (conditional_dm3d.py:186-195)."""
        B = x.shape[0]
        scale = float(c) ** -0.5
        n = emu.a(self._bn_apply(P, f"{name}.norm", x))
        h = emu.a(torch.relu(ops.conv3d(n, emu.w(P[f"{name}.proj_in.kernel"]), P[f"{name}.proj_in.bias"])))
        hf = h.reshape(B, s ** 3, c)
        ln = [emu.a(ops.layernorm(hf, P[f"{name}.norm{i}.gamma"], P[f"{name}.norm{i}.beta"])) for i in (1, 2, 3)]
        q, k, v = (emu.a(self._d(P, f"{name}.{m}", ln[0], emu)) for m in ("query", "key", "value"))
        t1 = emu.a(ops.attention_core(q, k, v, scale, emu) + hf)
        q2 = emu.a(self._d(P, f"{name}.query", ln[1], emu))
        kc, vc = self.context_kv(P, name, cemb, c, s, emu)
        t2 = emu.a(ops.attention_core(q2, kc, vc, scale, emu) + t1)
        m = emu.a(self._d(P, f"{name}.mlp0", ln[2], emu, act="relu"))
        xs = emu.a(self._d(P, f"{name}.mlp1", m, emu) + t2).reshape(x.shape)
        return emu.a(torch.relu(ops.conv3d(xs, emu.w(P[f"{name}.proj_out.kernel"]), P[f"{name}.proj_out.bias"])) + x)

    def forward(self, P, x, t, ctx=None, emu: Emu = EXACT):
        """x (B,S,S,S,C_lat) fp32, t (B,) int, ctx (B,) int class ids (conditional) -> eps_hat fp32."""
        temb = self.temb(P, t)
        cemb = P["ctx.embedding"][ctx.long()] if self.cond else None  # Embedding (conditional_dm3d.py:358)
        x = emu.a(ops.conv3d(emu.a(x), emu.w(P["in.kernel"]), P["in.bias"]), "in")
        skips = [x]
        for op in self.prog:
            kind = op[0]
            if kind == "res":
                x = self._resblock(P, op[1], x, temb, op[2], op[3], emu)
            elif kind == "attn":
                x = (self._xattention(P, op[1], x, cemb, op[2], op[3], emu) if self.cond
                     else self._attention(P, op[1], x, op[2], emu))
            elif kind == "push":
                skips.append(x)
            elif kind == "down":
                x = emu.a(ops.conv3d(x, emu.w(P[f"{op[1]}.kernel"]), P[f"{op[1]}.bias"], stride=2), op[1])
            elif kind == "cat":
                x = torch.cat([x, skips.pop()], dim=-1)
            elif kind == "up":
                x = emu.a(ops.conv3d(ops.upsample_nearest2(x), emu.w(P[f"{op[1]}.kernel"]), P[f"{op[1]}.bias"]), op[1])
        x = emu.a(ops.swish(self._bn_apply(P, "out.norm", x)), "out.norm")
        return ops.conv3d(x, emu.w(P["out.conv.kernel"]), P["out.conv.bias"])

    def flops(self, batch=1):
        """Algorithmic FLOPs (2*MACs) of one forward: convs, 1^3 convs / dense-on-voxels, attention core."""
        tot = dict(conv3=0, conv1=0, attn_proj=0, attn_core=0, mlp=0)
        s = self.S
        tot["conv3"] += 2 * 27 * self.C_lat * self.F * s ** 3
        for op in self.prog:
            if op[0] == "res":
                cin, w = op[2], op[3]
                if cin != w:
                    tot["conv1"] += 2 * cin * w * s ** 3
                tot["conv3"] += 2 * 27 * (cin * w + w * w) * s ** 3
            elif op[0] == "down":
                s //= 2
                c = self._c_of(op[1]); tot["conv3"] += 2 * 27 * c * c * s ** 3
            elif op[0] == "up":
                s *= 2
                c = self._c_of(op[1]); tot["conv3"] += 2 * 27 * c * c * s ** 3
            elif op[0] == "attn":
                c, L = op[2], op[3] ** 3
                if self.cond:
                    tot["conv1"] += 2 * 2 * c * c * L
                    tot["attn_proj"] += 2 * 4 * c * c * L
                    tot["attn_core"] += 2 * 4 * L * L * c
                    tot["mlp"] += 2 * 8 * c * c * L
                else:
                    tot["attn_proj"] += 2 * 4 * c * c * L
                    tot["attn_core"] += 4 * L * L * c
        tot["conv3"] += 2 * 27 * self.widths[0] * self.C_lat * self.S ** 3
        return {k: v * batch for k, v in tot.items()}

    def _c_of(self, name):
        for n, shp, _ in self._spec:
            if n == f"{name}.kernel":
                return shp[-1]
        raise KeyError(name)
