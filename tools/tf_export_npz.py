"""TensorFlow-side exporter: run this INSIDE the reference's environment (TensorFlow 2.12 + the reference repo on
PYTHONPATH).  It cannot run in this repository's container (no TensorFlow); it is the script INTEGRATION.md tells a
maintainer to run once per trained model.

It does two things:

1. weights -> a flat ``.npz`` of canonical names (``b200dm.weights`` layout): the reference model's layers are walked in
   their construction order and zipped onto ``b200dm.param_spec`` with shape checks, so the correspondence between Keras'
   ``model.network.layers`` and ``layer_with_weights-<n>`` checkpoint keys never has to be guessed;
2. (``--goldens``) golden vectors that close the "parity unpinned" gap: ``network([x, t(, ctx)])``, ``decoder(z)`` and
   ``quantizer.get_code_indices(z)`` evaluated by TensorFlow on seeded inputs, for ``tests/golden/``.

    python tools/tf_export_npz.py --ckpt checkpoints/dm3d-100 --out dm3d-100.npz [--conditional] [--goldens goldens.npz]
"""
import argparse
import sys

import numpy as np


def keras_layers_in_construction_order(model):
    """Keras assigns every layer a monotonically increasing uid suffix at construction (``conv3d_17``); sorting the
    weighted layers by that suffix recovers the order in which build_model created them (dm3d.py:294-376)."""
    def uid(layer):
        name = layer.name
        tail = name.rsplit("_", 1)[-1]
        return (name.rsplit("_", 1)[0], int(tail)) if tail.isdigit() else (name, 0)
    layers = [l for l in model.layers if l.weights]
    # global creation order = order of first weight creation; Keras names variables "<layer>/<attr>:0"
    return sorted(layers, key=lambda l: min(w._unique_id if hasattr(w, "_unique_id") else uid(l)[1] for w in l.weights))


def export_unet(network, spec):
    """{canonical name: ndarray} by zipping construction-ordered Keras layers onto ``spec`` (shape-checked)."""
    groups, cur = [], None
    for name, shape, _ in spec:
        stem, leaf = name.rsplit(".", 1)
        if leaf == "embedding":
            stem = name
        if cur is None or cur[0] != stem:
            cur = (stem, [])
            groups.append(cur)
        cur[1].append((name, leaf, tuple(shape)))
    layers = keras_layers_in_construction_order(network)
    if len(layers) != len(groups):
        sys.exit(f"model has {len(layers)} weighted layers, spec has {len(groups)} -- hyper-parameters differ")
    attr = {"kernel": "kernel", "bias": "bias", "gamma": "gamma", "beta": "beta", "mean": "moving_mean", "var": "moving_variance",
            "alpha": "alpha", "embedding": "embeddings"}
    out = {}
    for layer, (stem, tensors) in zip(layers, groups):
        by_attr = {w.name.rsplit("/", 1)[-1].split(":")[0]: w.numpy() for w in layer.weights}
        for name, leaf, shape in tensors:
            a = by_attr.get(attr[leaf])
            if a is None or tuple(a.shape) != shape:
                sys.exit(f"{layer.name}: no '{attr[leaf]}' of shape {shape} for {name} (layer holds {[(k, v.shape) for k, v in by_attr.items()]})")
            out[name] = a.astype(np.float32)
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
    ap.add_argument("--goldens", default=None)
    a = ap.parse_args()

    import types
    import tensorflow as tf                                     # noqa: F401  (reference environment only)
    import b200dm                                                # this repository, for param_spec / canonical names
    if a.conditional:
        from networks.conditional_dm3d import DiffusionModel    # reference
    else:
        from networks.dm3d import DiffusionModel
    args = types.SimpleNamespace(timesteps=a.timesteps, num_gpus=1, kernel_resize=False, bs=1, lr=1e-4)
    dm = DiffusionModel(a.latent_size, a.num_embed, a.latent_channels, None, args)
    dm.load_weights(a.ckpt).expect_partial()
    net = b200dm.build_model(a.latent_size, a.latent_channels, [64, 128, 256], [False, False, True, True],
                             context_dim=1 if a.conditional else None)
    params = export_unet(dm.network, net.spec)
    np.savez(a.out, **params)
    print(f"wrote {a.out}: {len(params)} tensors, {sum(v.size for v in params.values())} parameters")

    if a.goldens:
        rng = np.random.default_rng(1234)
        B, S, C = 2, a.latent_size, a.latent_channels
        x = rng.standard_normal((B, S, S, S, C)).astype(np.float32)
        t = np.array([17, 903], dtype=np.int64)
        ins = [x, t] + ([np.array([[[0]], [[1]]], dtype=np.int64)] if a.conditional else [])
        eps = dm.network(ins, training=False).numpy()
        np.savez(a.goldens, x=x, t=t, eps=eps)
        print(f"wrote {a.goldens}: network([x, t]) on seeded inputs (compare with tests/test_model_gpu.py tolerances)")


if __name__ == "__main__":
    main()
