"""TensorFlow-side exporter: run this INSIDE the reference's environment (TensorFlow 2.12 + the reference repo on
PYTHONPATH).  It cannot run in this repository's container (no TensorFlow); it is the script INTEGRATION.md tells a
maintainer to run once per trained model.  UNVALIDATED against a real checkpoint (none ships with the reference).

How tensors are matched (no dependence on Keras' ``model.layers`` order, which is DEPTH-sorted for functional models, nor on
``layer_with_weights-<n>`` numbering):

  * every LEAF layer that owns variables is collected recursively -- flat layers of the functional U-Net and the sublayers of
    custom blocks alike (``AttentionBlock.query``, ``CrossAttentionBlock.proj.layers[0]``, ``VQVAEResidualUnit.conv1`` ...);
  * leaves are bucketed by Keras class (Conv3D, Conv3DTranspose, Dense, BatchNormalization, LayerNormalization,
    GroupNormalization, PReLU, Embedding) and each bucket is sorted by the layer's auto-name counter (``dense_17``): Keras
    assigns that counter in ``Layer.__init__``, i.e. in the order the reference's code CONSTRUCTS layers of that class
    (dm3d.py:294-376 top to bottom; block sublayers in their ``__init__`` order);
  * this repository's canonical spec lists tensors in that same construction order, so walking the spec and taking "the next
    layer of the class this entry needs" reproduces the correspondence; every tensor is shape-checked, and a bucket that is
    not consumed exactly raises.

    python tools/tf_export_npz.py --ckpt checkpoints/dm3d-100 --out dm3d-100.npz [--conditional]
           [--first-stage-ckpt checkpoints/vqvae-50 --first-stage-out vqvae-50.npz] [--goldens goldens.npz]

``--goldens`` writes TensorFlow-evaluated ``network([x, t(, ctx)])`` / ``decoder(z)`` / ``get_code_indices(z)`` on seeded inputs:
the vectors that would close this repository's "parity unpinned" status (tests/golden/)."""
import argparse
import re
import sys

import numpy as np

LEAF_ATTR = {"kernel": "kernel", "bias": "bias", "gamma": "gamma", "beta": "beta", "mean": "moving_mean", "var": "moving_variance",
             "alpha": "alpha", "embedding": "embeddings"}


def spec_groups(spec, gamma_beta_class="LayerNormalization", up_class="Conv3D"):
    """[(stem, keras class bucket, [(canonical name, leaf, shape)])] in spec (= construction) order.
    ``gamma_beta_class``: class of norms that hold only (gamma, beta): LayerNormalization in the U-Net's CrossAttentionBlock,
    GroupNormalization in the vqgan_gnorm / vqgan_attn_cp first stages.  ``up_class``: class of ``*.up.kernel`` (k = 4) layers:
    Conv3DTranspose in every first-stage decoder except vqgan_stride (Conv3D + UpSampling3D)."""
    groups, cur = [], None
    for name, shape, _ in spec:
        stem, leaf = name.rsplit(".", 1) if "." in name else ("", name)
        if leaf == "embedding":
            stem = name
        if cur is None or cur[0] != stem:
            cur = [stem, None, []]
            groups.append(cur)
        cur[2].append((name, leaf, tuple(shape)))
    for g in groups:
        leaves = {leaf: shape for _, leaf, shape in g[2]}
        if "embedding" in leaves:
            g[1] = "Embedding"
        elif "alpha" in leaves:
            g[1] = "PReLU"
        elif "mean" in leaves:
            g[1] = "BatchNormalization"
        elif "gamma" in leaves:
            g[1] = gamma_beta_class
        elif len(leaves["kernel"]) == 5:
            g[1] = up_class if g[0].endswith(".up") else "Conv3D"
        else:
            g[1] = "Dense"
    return groups


def leaf_layers(model):
    """{class name: [leaf layers with variables, sorted by their auto-name counter]} over the model and all nested sublayers."""
    def counter(layer):
        m = re.search(r"_(\d+)$", layer.name)
        return int(m.group(1)) if m else 0
    buckets = {}
    for layer in model._flatten_layers(include_self=False, recursive=True):
        if not layer.weights or any(True for _ in layer._flatten_layers(include_self=False, recursive=False)):
            continue   # unbuilt (AttentionBlock.depth is never called) or a container
        buckets.setdefault(type(layer).__name__, []).append(layer)
    return {k: sorted(v, key=counter) for k, v in buckets.items()}


def export_by_class(model, spec, what, **classes):
    buckets, used, out = leaf_layers(model), {}, {}
    for stem, cls, tensors in spec_groups(spec, **classes):
        seq = buckets.get(cls, [])
        i = used.get(cls, 0)
        if i >= len(seq):
            sys.exit(f"{what}: the reference model has only {len(seq)} {cls} layers with weights, the spec needs more (at {stem})")
        layer, used[cls] = seq[i], i + 1
        by_attr = {w.name.rsplit("/", 1)[-1].split(":")[0]: w.numpy() for w in layer.weights}
        for name, leaf, shape in tensors:
            a = by_attr.get(LEAF_ATTR[leaf])
            if a is None or tuple(a.shape) != shape:
                sys.exit(f"{what}: {layer.name} ({cls} #{i}) has no '{LEAF_ATTR[leaf]}' of shape {shape} for {name}; it holds "
                         f"{[(k, v.shape) for k, v in by_attr.items()]}")
            out[name] = a.astype(np.float32)
    for cls, seq in buckets.items():
        if used.get(cls, 0) != len(seq):
            sys.exit(f"{what}: {len(seq) - used.get(cls, 0)} {cls} layers of the reference model were not consumed -- hyper-parameters differ")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--conditional", action="store_true")
    ap.add_argument("--latent-size", type=int, default=32)
    ap.add_argument("--latent-channels", type=int, default=256)
    ap.add_argument("--num-embed", type=int, default=1024)
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--first-stage-ckpt", default=None, help="checkpoint of the first-stage trainer (vqvae_trainer.load_weights)")
    ap.add_argument("--first-stage-out", default=None)
    ap.add_argument("--goldens", default=None)
    a = ap.parse_args()

    import types
    import tensorflow as tf                                     # noqa: F401  (reference environment only)
    import b200dm                                                # this repository, for the canonical specs
    if a.conditional:
        from networks.conditional_dm3d import DiffusionModel    # reference
    else:
        from networks.dm3d import DiffusionModel
    args = types.SimpleNamespace(timesteps=a.timesteps, num_gpus=1, kernel_resize=False, bs=1, lr=1e-4)
    dm = DiffusionModel(a.latent_size, a.num_embed, a.latent_channels, a.first_stage_ckpt, args)
    dm.load_weights(a.ckpt).expect_partial()
    net = b200dm.build_model(a.latent_size, a.latent_channels, [64, 128, 256], [False, False, True, True],
                             context_dim=1 if a.conditional else None)
    params = export_by_class(dm.network, net.spec, "U-Net")
    np.savez(a.out, **params)
    print(f"wrote {a.out}: {len(params)} tensors, {sum(v.size for v in params.values())} parameters")

    fs = None
    if a.first_stage_out:
        # the DM's first stage (dm3d.py:386-404 / conditional_dm3d.py:436-452): monai VQ-VAE, channels (32,64,128,256), R=5
        ch = (32, 64, 128, 256)
        fs = b200dm.VQVAE(1, 1, ch, 5, ch, num_embeddings=a.num_embed, embedding_dim=a.latent_channels, latent_size=a.latent_size)
        out = {"quantizer.embeddings": dm.quantizer.embeddings.numpy().astype(np.float32)}
        out.update({"decoder." + k: v for k, v in export_by_class(dm.decoder, fs.decoder.spec, "decoder", up_class="Conv3DTranspose").items()})
        out.update({"encoder." + k: v for k, v in export_by_class(dm.encoder, fs.encoder.spec, "encoder").items()})
        np.savez(a.first_stage_out, **out)
        print(f"wrote {a.first_stage_out}: {len(out)} tensors")

    if a.goldens:
        rng = np.random.default_rng(1234)
        B, S, C = 2, a.latent_size, a.latent_channels
        x = rng.standard_normal((B, S, S, S, C)).astype(np.float32)
        t = np.array([17, 903], dtype=np.int64)
        ins = [x, t] + ([np.array([[[0]], [[1]]], dtype=np.int64)] if a.conditional else [])
        g = dict(x=x, t=t, eps=dm.network(ins, training=False).numpy())
        z = (rng.standard_normal((1, a.latent_size, a.latent_size, a.latent_size, C)) * 0.5).astype(np.float32)
        g["z"] = z
        g["code_indices"] = dm.quantizer.get_code_indices(z.reshape(-1, C)).numpy()
        g["decoded"] = dm.decoder(z, training=False).numpy()
        np.savez(a.goldens, **g)
        print(f"wrote {a.goldens}: network([x, t(, ctx)]), get_code_indices(z), decoder(z) evaluated by TensorFlow on seeded inputs")


if __name__ == "__main__":
    main()
