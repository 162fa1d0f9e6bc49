"""TensorFlow tensor-bundle checkpoints: format restatement pinned by published known answers, a writer ->
reader round trip, and the object-graph-key -> canonical-name assignment for the U-Net.  TensorFlow itself is absent."""
import os
import struct

import numpy as np
import pytest

import b200dm
from b200dm import tf_checkpoint as T


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors for CRC32C (Castagnoli)
    assert T.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(b"123456789") == 0xE3069283
    # tensorflow/core/lib/hash/crc32c.h: Mask(crc) = rotr(crc, 15) + 0xa282ead8
    assert T.mask_crc(0) == 0xA282EAD8


def test_varint_and_proto_roundtrip():
    for v in (0, 1, 127, 128, 300, 2 ** 31, 2 ** 40 + 5):
        assert T._get_varint(T._put_varint(v), 0) == (v, len(T._put_varint(v)))
    msg = T._field(1, 0, T._put_varint(1)) + T._field(4, 0, T._put_varint(1 << 33)) + T._field(6, 5, struct.pack("<I", 0xDEADBEEF))
    m = T._parse_proto(msg)
    assert m[1] == [1] and m[4] == [1 << 33] and m[6] == [0xDEADBEEF]


def test_write_read_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    # enough keys that the index spans several data blocks and exercises key prefix compression
    vars_ = {f"network/layer_with_weights-{i}/kernel/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal((3, 3, 3, 4, 8)).astype(np.float32)
             for i in range(60)}
    vars_.update({f"network/layer_with_weights-{i}/bias/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal((8,)).astype(np.float32) for i in range(60)})
    vars_["step/.ATTRIBUTES/VARIABLE_VALUE"] = np.array(7, dtype=np.int64)
    prefix = str(tmp_path / "ckpt" / "dm-3")
    T.write_checkpoint(prefix, vars_)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == T.MAGIC            # LevelDB table magic (table/format.h kTableMagicNumber)
    header, entries = T.read_index(prefix + ".index")
    assert header == dict(num_shards=1, endianness=0) and len(entries) == len(vars_)
    got = T.read_checkpoint(prefix, verify=True)
    assert set(got) == set(vars_)
    for k, v in vars_.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "c")
    T.write_checkpoint(prefix, {"a/.ATTRIBUTES/VARIABLE_VALUE": np.arange(8, dtype=np.float32)})
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[3] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        T.read_index(prefix + ".index")
    with pytest.raises(ValueError):
        T.read_index(prefix + ".data-00000-of-00001")   # not a table at all


def test_unet_weights_from_object_graph_checkpoint(tmp_path):
    """Save a small U-Net's weights under Keras object-graph keys, reload through the canonical names."""
    net = b200dm.build_model(8, 8, [64, 128, 256], [False, False, True, True])
    params = net.params
    # one checkpoint layer per weighted Keras layer, numbered in the construction order of param_spec
    layers, cur = [], None
    for name, shape, _ in net.spec:
        stem, leaf = name.rsplit(".", 1)
        if cur is None or cur[0] != stem:
            cur = (stem, [])
            layers.append(cur)
        cur[1].append((name, leaf))
    vars_ = {}
    for n, (stem, tensors) in enumerate(layers):
        for name, leaf in tensors:
            attr = T._ATTR_OF_LEAF.get(leaf, leaf)
            vars_[f"network/layer_with_weights-{n}/{attr}/.ATTRIBUTES/VARIABLE_VALUE"] = params[name].numpy()
    vars_["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"] = np.array(3, dtype=np.int64)
    prefix = str(tmp_path / "dm3d-1")
    T.write_checkpoint(prefix, vars_, checksum_limit=1 << 14)
    # a bare TF checkpoint of the functional U-Net does not determine the layer of each layer_with_weights-<n> (Keras sorts
    # model.layers by depth): refused by default, zipped by n only on request (this file was written in creation order)
    with pytest.raises(ValueError, match="tf_export_npz"):
        T.load_keras_checkpoint(prefix, net.spec, root="network")
    loaded = T.load_keras_checkpoint(prefix, net.spec, root="network", assume_creation_order=True)
    assert set(loaded) == set(params)
    for k in params:
        assert np.array_equal(loaded[k], params[k].numpy()), k
    net2 = b200dm.build_model(8, 8, [64, 128, 256], [False, False, True, True])
    net2.load_weights(prefix, assume_creation_order=True)
    assert all(np.array_equal(net2.params[k].numpy(), params[k].numpy()) for k in params)
    # the reliable route: an explicit {canonical name: checkpoint key} map (what the TF-side exporter knows)
    name_map = {name: f"network/layer_with_weights-{n}/{T._ATTR_OF_LEAF.get(leaf, leaf)}/.ATTRIBUTES/VARIABLE_VALUE"
                for n, (stem, tensors) in enumerate(layers) for name, leaf in tensors}
    net4 = b200dm.build_model(8, 8, [64, 128, 256], [False, False, True, True])
    net4.load_weights(prefix, name_map=name_map)
    assert all(np.array_equal(net4.params[k].numpy(), params[k].numpy()) for k in params)
    # nested block variables (AttentionBlock is one Keras layer) are recognised and never zipped by order
    vars_n = dict(vars_)
    vars_n["network/layer_with_weights-900/query/kernel/.ATTRIBUTES/VARIABLE_VALUE"] = np.zeros((4, 4), np.float32)
    T.write_checkpoint(str(tmp_path / "nested"), vars_n, checksum_limit=1 << 14)
    assert (900, "query/kernel") in T.keras_nested_variables(T.read_checkpoint(str(tmp_path / "nested")), "network")
    with pytest.raises(ValueError, match="nested"):
        T.load_keras_checkpoint(str(tmp_path / "nested"), net.spec, assume_creation_order=True)
    # a checkpoint of a different architecture is rejected with the offending shapes
    net3 = b200dm.build_model(8, 16, [64, 128, 256], [False, False, True, True])
    with pytest.raises(ValueError):
        T.load_keras_checkpoint(prefix, net3.spec, assume_creation_order=True)


def test_first_stage_weights_from_checkpoint(tmp_path):
    """vqvae_trainer.load_weights(ckpt): decoder blocks (Sequential order) + quantizer codebook from object-graph keys."""
    vq = b200dm.VQVAE(1, 1, (32, 64), 1, (32, 64), num_embeddings=16, embedding_dim=8, latent_size=4)
    spec = vq.decoder.spec
    rng = np.random.default_rng(1)
    layers, cur = [], None
    for name, shape, _ in spec:
        stem, leaf = name.rsplit(".", 1)
        if cur is None or cur[0] != stem:
            cur = (stem, [])
            layers.append(cur)
        cur[1].append((name, leaf, shape))
    vars_, want = {}, {}
    for n, (stem, tensors) in enumerate(layers):
        for name, leaf, shape in tensors:
            a = rng.standard_normal(shape).astype(np.float32)
            want[name] = a
            vars_[f"decoder/blocks/layer_with_weights-{n}/{T._ATTR_OF_LEAF.get(leaf, leaf)}/.ATTRIBUTES/VARIABLE_VALUE"] = a
    cb = rng.standard_normal((8, 16)).astype(np.float32)           # monai codebook layout (D, K)
    vars_["quantizer/embeddings/.ATTRIBUTES/VARIABLE_VALUE"] = cb
    vars_["quantizer/codebooks_used/.ATTRIBUTES/VARIABLE_VALUE"] = np.zeros(16, dtype=np.int32)
    prefix = str(tmp_path / "vqvae-5")
    T.write_checkpoint(prefix, vars_, checksum_limit=1 << 14)
    vq.load_weights(prefix)
    for k, a in want.items():
        assert np.array_equal(vq.decoder.params[k].numpy(), a), k
    assert np.array_equal(np.asarray(vq.quantizer.embeddings), cb)


def test_first_stage_nested_object_graph_keys_and_encoder(tmp_path):
    """The monai first stage as the reference's object graph tracks it (vqvae3d_monai.py:218-234, 253-306, 342-391):
    ``blocks`` is a Sequential; a VQVAEResidualUnit is a nested Model whose ``conv1`` is a layer attribute and whose ``conv2`` is
    an inner Sequential (Conv3D, BatchNormalization, PReLU) -> keys ``.../layer_with_weights-N/conv2/layer_with_weights-M/...``.
    vqvae_trainer.load_weights restores encoder, quantizer AND decoder: both are filled here, and a checkpoint
    without encoder tensors makes encode() refuse instead of running on random weights."""
    vq = b200dm.VQVAE(1, 1, (32, 64), 1, (32, 64), num_embeddings=16, embedding_dim=8, latent_size=4)
    rng = np.random.default_rng(2)
    vars_, want = {}, {"decoder": {}, "encoder": {}}
    for part, spec in (("decoder", vq.decoder.spec), ("encoder", vq.encoder.spec)):
        n = -1          # Sequential index of the current weighted block
        last_unit = None
        for name, shape, _ in spec:
            a = rng.standard_normal(shape).astype(np.float32)
            want[part][name] = a
            stem, leaf = name.rsplit(".", 1)
            attr = T._ATTR_OF_LEAF.get(leaf, leaf)
            if ".res." in stem:
                unit, sub = stem.rsplit(".", 1)          # level.i.res.j , conv1|conv2|norm|prelu
                if unit != last_unit:
                    n += 1
                    last_unit = unit
                inner = {"conv1": "conv1", "conv2": "conv2/layer_with_weights-0", "norm": "conv2/layer_with_weights-1",
                         "prelu": "conv2/layer_with_weights-2"}[sub]
                key = f"{part}/blocks/layer_with_weights-{n}/{inner}/{attr}"
            else:
                if leaf in ("kernel", "alpha") or last_unit is not None:
                    if leaf != "bias":
                        n += 1
                    last_unit = None
                key = f"{part}/blocks/layer_with_weights-{n}/{attr}"
            vars_[key + "/.ATTRIBUTES/VARIABLE_VALUE"] = a
    cb = rng.standard_normal((8, 16)).astype(np.float32)
    vars_["quantizer/embeddings/.ATTRIBUTES/VARIABLE_VALUE"] = cb
    prefix = str(tmp_path / "vqvae-nested")
    T.write_checkpoint(prefix, vars_, checksum_limit=1 << 14)
    vq.load_weights(prefix)
    for part, obj in (("decoder", vq.decoder), ("encoder", vq.encoder)):
        for k, a in want[part].items():
            assert np.array_equal(obj.params[k].numpy(), a), (part, k)
    assert vq.encoder.weights_missing is None
    # decoder-only checkpoint: loads, but the encoder refuses to run
    dec_only = {k: v for k, v in vars_.items() if not k.startswith("encoder/")}
    prefix2 = str(tmp_path / "vqvae-deconly")
    T.write_checkpoint(prefix2, dec_only, checksum_limit=1 << 14)
    vq2 = b200dm.VQVAE(1, 1, (32, 64), 1, (32, 64), num_embeddings=16, embedding_dim=8, latent_size=4)
    vq2.load_weights(prefix2)
    assert vq2.encoder.weights_missing == prefix2
    import torch
    with pytest.raises(b200dm._lib.B200dmError):
        vq2.encoder(torch.zeros(1, 16, 16, 16, 1))


def test_first_stage_npz_roundtrip_includes_encoder(tmp_path):
    vq = b200dm.VQVAE(1, 1, (32, 64), 1, (32, 64), num_embeddings=16, embedding_dim=8, latent_size=4)
    rng = np.random.default_rng(3)
    flat = {f"decoder.{n}": rng.standard_normal(s).astype(np.float32) for n, s, _ in vq.decoder.spec}
    flat.update({f"encoder.{n}": rng.standard_normal(s).astype(np.float32) for n, s, _ in vq.encoder.spec})
    flat["quantizer.embeddings"] = rng.standard_normal((8, 16)).astype(np.float32)
    path = str(tmp_path / "fs.npz")
    np.savez(path, **flat)
    vq.load_weights(path)
    for n, _, _ in vq.encoder.spec:
        assert np.array_equal(vq.encoder.params[n].numpy(), flat[f"encoder.{n}"]), n
    assert vq.encoder.weights_missing is None


def test_tf_side_exporter_matches_by_class_creation_order_not_by_layer_order():
    """tools/tf_export_npz.py (runs inside the reference's TensorFlow environment) pairs tensors by Keras class + auto-name
    counter, so neither the DEPTH-sorted ``model.layers`` of a functional model nor nesting inside custom blocks matters.
    Mock of the Keras surface it touches: layers named ``<class>_<n>`` in construction order, some nested in container layers,
    handed over in a shuffled order (zip-by-order exporters silently mispair)."""
    import importlib.util
    import random
    import b200dm
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec_ = importlib.util.spec_from_file_location("tf_export_npz", os.path.join(ROOT, "tools", "tf_export_npz.py"))
    E = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(E)

    class W:
        def __init__(self, name, a):
            self.name, self.a = name, a

        def numpy(self):
            return self.a

    class Layer:
        def __init__(self, name, weights=(), children=()):
            self.name, self.weights, self.children = name, list(weights), list(children)

        def _flatten_layers(self, include_self=False, recursive=True):
            for c in self.children:
                yield c
                if recursive:
                    yield from c._flatten_layers(False, True)

    net = b200dm.build_model(8, 8, [64, 128, 256], [False, False, True, True], context_dim=1)
    rng = np.random.default_rng(0)
    snake = {"Conv3D": "conv3d", "Dense": "dense", "BatchNormalization": "batch_normalization", "LayerNormalization": "layer_normalization",
             "Embedding": "embedding"}
    counters, flat, want = {}, [], {}
    for stem, cls, tensors in E.spec_groups(net.spec):
        i = counters.get(cls, 0)
        counters[cls] = i + 1
        ws = []
        for name, leaf, shape in tensors:
            a = rng.standard_normal(shape).astype(np.float32)
            want[name] = a
            ws.append(W(f"{snake[cls]}_{i + 7}/{E.LEAF_ATTR[leaf]}:0", a))
        layer = type(cls, (Layer,), {})(f"{snake[cls]}_{i + 7}", ws)   # counters start at 7: another model was built first
        flat.append((stem, layer))
    # nest the sublayers of every attention block inside a container (a custom Keras Layer), then shuffle everything
    top, blocks = [], {}
    for stem, layer in flat:
        if ".attn." in stem and not stem.endswith("ctxmlp"):
            blocks.setdefault(stem.rsplit(".", 1)[0], []).append(layer)
        else:
            top.append(layer)
    for k, (bname, subs) in enumerate(blocks.items()):
        random.Random(k).shuffle(subs)
        top.append(Layer(f"cross_attention_block_{k}", weights=[w for s in subs for w in s.weights], children=subs))
    random.Random(1).shuffle(top)
    got = E.export_by_class(Layer("model", children=top), net.spec, "U-Net")
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
