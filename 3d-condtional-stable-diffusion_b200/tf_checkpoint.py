"""Pure-Python reader (and a minimal writer) for TensorFlow checkpoints in the tensor-bundle format.

The reference saves its models with ``ModelCheckpoint(save_weights_only=True)`` / ``model.save_weights(prefix)``
(main.py:392-396; reloaded by ``load_weights`` in dm3d.py:408-414 and main.py:440-446).  That writes

    <prefix>.index                   an SSTable (LevelDB table format): key -> serialized BundleEntryProto
    <prefix>.data-00000-of-00001     the raw tensor bytes, little-endian, at BundleEntryProto.offset/.size

TensorFlow is not installed here, so this module restates the published on-disk format (tensorflow/core/lib/io/
{format,block,table}.cc, tensorflow/core/util/tensor_bundle/tensor_bundle.cc, tensor_bundle.proto) with nothing but
``struct`` and numpy:

    table      [data blocks][metaindex block][index block][footer(48 B): metaindex handle, index handle, magic]
    block      entries (varint shared, varint non_shared, varint value_len, key suffix, value) + uint32 restarts[] + count;
               each block is followed by 1 byte compression type (0 = none) and a masked CRC32C
    key ""     BundleHeaderProto {num_shards, endianness, version}
    other keys BundleEntryProto {1 dtype, 2 TensorShapeProto, 3 shard_id, 4 offset, 5 size, 6 masked crc32c}

Keys of a Keras object-graph checkpoint look like ``network/layer_with_weights-7/kernel/.ATTRIBUTES/VARIABLE_VALUE``.
``read_checkpoint`` returns every numeric variable under its key; ``keras_layer_variables`` groups them per
``layer_with_weights-<n>``; ``assign_by_creation_order`` zips those groups onto a model's ``param_spec`` (the reference's
layer construction order) with shape checks.  The n -> layer correspondence of a *functional* Keras model follows
``model.layers`` (depth-sorted); without TensorFlow or a real checkpoint it cannot be confirmed here, so a mismatch raises
with the offending shapes and an explicit ``name_map`` can be supplied instead.  The writer exists for round-trip tests
and to hand weights back in a file layout TF tooling can list (``tf.train.list_variables``).
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           14: None, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}  # 14 = bfloat16 (returned as uint16 view -> float32)
_DT_OF = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9, np.dtype(np.bool_): 10,
          np.dtype(np.float16): 19}


# ------------------------------------------------------------------------------------------------ CRC32C (Castagnoli)
def _make_crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TAB = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TAB
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c: int) -> int:   # tensorflow/core/lib/hash/crc32c.h
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varint / protobuf
def _get_varint(buf, pos):
    r, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if not b & 0x80:
            return r, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """{field: [values]} of one message; wire types 0 (varint), 1 (64-bit), 2 (bytes), 5 (32-bit)."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n]); pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(f, []).append(v)
    return out


def _field(tag, wt, payload):
    return _put_varint((tag << 3) | wt) + payload


# ------------------------------------------------------------------------------------------------ table (SSTable) reader
def _read_block(buf, offset, size, verify=True):
    data = buf[offset:offset + size]
    ctype = buf[offset + size]
    if verify:
        want = struct.unpack_from("<I", buf, offset + size + 1)[0]
        got = mask_crc(crc32c(bytes(buf[offset:offset + size + 1])))
        if want != got:
            raise ValueError(f"checkpoint index: block at {offset} fails its CRC32C")
    if ctype != 0:
        raise NotImplementedError("checkpoint index block is Snappy-compressed; TensorFlow writes bundle indexes uncompressed")
    return data


def _block_entries(block):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared]); pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(index_path, verify=True):
    """-> (header dict, {key: entry dict}) of a ``.index`` file."""
    buf = memoryview(open(index_path, "rb").read())
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError(f"{index_path}: not a TensorFlow tensor-bundle index (bad magic)")
    foot = buf[len(buf) - 48:]
    _, p = _get_varint(foot, 0)          # metaindex handle (unused)
    _, p = _get_varint(foot, p)
    ioff, p = _get_varint(foot, p)
    isize, p = _get_varint(foot, p)
    header, entries = None, {}
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, q = _get_varint(handle, 0)
        bsize, q = _get_varint(handle, q)
        for key, val in _block_entries(_read_block(buf, boff, bsize, verify)):
            m = _parse_proto(val)
            if key == b"":
                header = dict(num_shards=m.get(1, [1])[0], endianness=m.get(2, [0])[0])
                continue
            shape = []
            if 2 in m:
                for dim in _parse_proto(m[2][0]).get(2, []):
                    shape.append(_parse_proto(dim).get(1, [0])[0])
            entries[key.decode()] = dict(dtype=m.get(1, [0])[0], shape=tuple(shape), shard=m.get(3, [0])[0], offset=m.get(4, [0])[0],
                                         size=m.get(5, [0])[0], crc=m.get(6, [None])[0], sliced=7 in m)
    if header is None:
        raise ValueError(f"{index_path}: no bundle header entry")
    if header["endianness"] != 0:
        raise NotImplementedError("big-endian tensor bundle")
    return header, entries


def read_checkpoint(prefix, verify=False, keys=None):
    """{key: ndarray} of every numeric variable in the checkpoint ``prefix`` (``prefix.index`` + ``prefix.data-*``).
    ``verify`` checks every tensor's CRC32C (slow in pure Python: ~1 s per MB)."""
    header, entries = read_index(prefix + ".index")
    shards = {}
    out = {}
    for key, e in entries.items():
        if keys is not None and key not in keys:
            continue
        if e["dtype"] not in _DTYPES or e["sliced"]:
            continue   # strings (the object graph), resources, partitioned variables
        if e["shard"] not in shards:
            shards[e["shard"]] = np.memmap(f"{prefix}.data-{e['shard']:05d}-of-{header['num_shards']:05d}", dtype=np.uint8, mode="r")
        raw = shards[e["shard"]][e["offset"]:e["offset"] + e["size"]]
        if verify and e["crc"] is not None and mask_crc(crc32c(raw.tobytes())) != e["crc"]:
            raise ValueError(f"{key}: tensor bytes fail their CRC32C")
        if e["dtype"] == 14:   # bfloat16 -> float32
            a = (np.frombuffer(raw.tobytes(), dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)
        else:
            a = np.frombuffer(raw.tobytes(), dtype=_DTYPES[e["dtype"]])
        out[key] = a.reshape(e["shape"]).copy()
    return out


# ------------------------------------------------------------------------------------------------ Keras object-graph keys
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def keras_layer_variables(variables, root):
    """Group ``{key: array}`` under ``<root>/layer_with_weights-<n>/<attr>`` -> [(n, {attr: array})] sorted by n.
    Optimizer slots (``.OPTIMIZER_SLOT``) and non-layer keys are ignored."""
    pat = re.compile(re.escape(root.rstrip("/")) + r"/layer_with_weights-(\d+)/([A-Za-z_0-9]+)" + re.escape(_SUFFIX) + r"$")
    groups = {}
    for k, v in variables.items():
        m = pat.match(k)
        if m:
            groups.setdefault(int(m.group(1)), {})[m.group(2)] = v
    return sorted(groups.items())


def block_layer_variables(variables, root):
    """Layer groups of a first-stage encoder / decoder under ``<root>`` = ``decoder/blocks`` or ``encoder/blocks``, in block
    order.  Handles the three ways the reference's blocks are tracked:
      * a Sequential:  ``<root>/layer_with_weights-<n>/<attr>``
      * a Python list: ``<root>/<i>/<attr>``  (a plain layer at list index i)
      * nested models (VQVAEResidualUnit: conv1, conv2[, norm], PReLU as attributes / inner Sequentials):
        ``<root>/<i>/<sub>/<attr>`` or ``<root>/<i>/<sub>/layer_with_weights-<m>/<attr>`` -- flattened to one group per inner
        layer, ordered (i, position of <sub> in the unit's construction order, m).
    Returns [(index tuple, {attr: array})] sorted; every assignment downstream is shape-checked."""
    root = root.rstrip("/") + "/"
    sub_order = {"conv1": 0, "conv2": 1, "norm": 2, "bn": 2, "act": 3, "prelu": 3}
    groups = {}
    for k, v in variables.items():
        if not (k.startswith(root) and k.endswith(_SUFFIX)) or ".OPTIMIZER_SLOT" in k:
            continue
        parts = k[len(root):-len(_SUFFIX)].split("/")
        attr, path = parts[-1], parts[:-1]
        key = []
        for comp in path:
            m = re.match(r"layer_with_weights-(\d+)$", comp)
            if m:
                key.append(int(m.group(1)))
            elif comp.isdigit():
                key.append(int(comp))
            else:
                key.append(sub_order.get(comp, 9))
        groups.setdefault(tuple(key), {})[attr] = v
    return sorted(groups.items())


_ATTR_OF_LEAF = {"kernel": "kernel", "bias": "bias", "gamma": "gamma", "beta": "beta", "mean": "moving_mean", "var": "moving_variance",
                 "alpha": "alpha", "embedding": "embeddings"}


def assign_by_creation_order(spec, layer_groups):
    """Zip checkpoint layer groups (``keras_layer_variables``) onto ``spec`` = [(canonical name, shape, init)] in the
    reference's layer construction order.  Consecutive spec entries sharing a name stem form one layer.  Every shape is
    checked; a mismatch raises with both sides so the correspondence can be fixed with an explicit map."""
    layers, cur = [], None
    for name, shape, _ in spec:
        stem, leaf = name.rsplit(".", 1)
        if leaf == "embedding":
            stem = name
        if cur is None or cur[0] != stem:
            cur = (stem, [])
            layers.append(cur)
        cur[1].append((name, leaf, tuple(shape)))
    if len(layers) != len(layer_groups):
        raise ValueError(f"checkpoint has {len(layer_groups)} layers with weights, the model has {len(layers)}")
    out = {}
    for (stem, tensors), (n, attrs) in zip(layers, layer_groups):
        for name, leaf, shape in tensors:
            attr = _ATTR_OF_LEAF.get(leaf, leaf)
            if attr not in attrs:
                raise KeyError(f"layer_with_weights-{n} has no '{attr}' (wanted for {name}); it holds {sorted(attrs)}")
            a = attrs[attr]
            if tuple(a.shape) != shape:
                raise ValueError(f"layer_with_weights-{n}/{attr} has shape {tuple(a.shape)}, {name} expects {shape}")
            out[name] = a.astype(np.float32)
    return out


def keras_nested_variables(variables, root):
    """{(n, 'sub/attr' path): array} for every ``<root>/layer_with_weights-<n>/<...>/<attr>`` key, nested sublayers included
    (AttentionBlock / CrossAttentionBlock are single Keras layers whose variables sit one or two levels down:
    ``layer_with_weights-31/query/kernel``, ``.../proj/layer_with_weights-0/kernel``)."""
    pat = re.compile(re.escape(root.rstrip("/")) + r"/layer_with_weights-(\d+)/(.+)" + re.escape(_SUFFIX) + r"$")
    out = {}
    for k, v in variables.items():
        m = pat.match(k)
        if m and ".OPTIMIZER_SLOT" not in k:
            out[(int(m.group(1)), m.group(2))] = v
    return out


def load_keras_checkpoint(prefix, spec, root="network", name_map=None, assume_creation_order=False):
    """{canonical name: float32 array} for a model with ``spec`` from the TF checkpoint ``prefix``.

    ``name_map`` {canonical name: checkpoint key} is the reliable route (tools/tf_export_npz.py, run where TensorFlow is,
    pairs tensors by Keras class + construction counter and can also write this map).  Without it the only information in the
    file is ``layer_with_weights-<n>``, and for a FUNCTIONAL model (the U-Net) n follows Keras' depth-sorted ``model.layers``
    -- norm1 sorts before the time-embedding Dense, a block's shortcut conv next to its conv2 -- not the order build_model
    constructs layers in, and attention blocks nest their variables.  Zipping by n is therefore only done on request
    (``assume_creation_order=True``: Sequential-style checkpoints, this repository's own writer) and is shape-checked; the
    default refuses instead of guessing.  Not validated against a checkpoint written by TensorFlow (none ships with the
    reference)."""
    variables = read_checkpoint(prefix)
    if name_map is not None:
        out = {}
        for name, shape, _ in spec:
            a = variables[name_map[name]]
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{name_map[name]} has shape {tuple(a.shape)}, {name} expects {tuple(shape)}")
            out[name] = a.astype(np.float32)
        return out
    nested = [k for k in keras_nested_variables(variables, root) if "/" in k[1]]
    if nested or not assume_creation_order:
        raise ValueError(
            f"{prefix}: the U-Net's variables are keyed by Keras' depth-sorted layer index (layer_with_weights-<n>"
            + (f", {len(nested)} of them nested inside block layers" if nested else "") + "), which does not determine the layer without "
            "TensorFlow.  Export canonical names with tools/tf_export_npz.py (inside the reference's environment) and load the .npz, "
            "or pass name_map={canonical name: checkpoint key}; assume_creation_order=True zips flat keys by n (shape-checked).")
    return assign_by_creation_order(spec, keras_layer_variables(variables, root))


# ------------------------------------------------------------------------------------------------ writer (tests, export)
def _block_bytes(entries, restart_interval=16):
    out, restarts, last = bytearray(), [], b""
    for i, (key, val) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(key), len(last)) and key[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_checkpoint(prefix, variables, checksum_limit=None):
    """Write ``{key: ndarray}`` as ``prefix.index`` + ``prefix.data-00000-of-00001`` (one shard, uncompressed).
    ``checksum_limit``: tensors with more bytes than this get no CRC32C field (the pure-Python CRC runs at a few MB/s;
    TensorFlow's reader insists on the checksum, so leave it at None for files TF has to open)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = sorted(((k.encode(), np.asarray(v)) for k, v in variables.items()), key=lambda kv: kv[0])
    data = bytearray()
    entries = [(b"", _field(1, 0, _put_varint(1)) + _field(2, 0, _put_varint(0)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1))))]
    for key, a in items:
        if a.dtype not in _DT_OF:
            raise TypeError(f"{key.decode()}: dtype {a.dtype} not supported by the writer")
        raw = a.tobytes()
        shape = b"".join(_field(2, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(int(s))))) for s in a.shape)
        e = _field(1, 0, _put_varint(_DT_OF[a.dtype])) + _field(2, 2, _put_varint(len(shape)) + shape)
        if len(data):
            e += _field(4, 0, _put_varint(len(data)))
        e += _field(5, 0, _put_varint(len(raw)))
        if checksum_limit is None or len(raw) <= checksum_limit:
            e += _field(6, 5, struct.pack("<I", mask_crc(crc32c(raw))))
        entries.append((key, e))
        data += raw
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    out = bytearray()

    def add_block(content):
        off = len(out)
        out.extend(content)
        out.append(0)
        out.extend(struct.pack("<I", mask_crc(crc32c(content + b"\x00"))))
        return _put_varint(off) + _put_varint(len(content))

    handles, chunk = [], []
    size = 0
    for key, val in entries:   # ~4 KB data blocks, like the table builder
        chunk.append((key, val))
        size += len(key) + len(val)
        if size >= 4096:
            handles.append((chunk[-1][0], add_block(_block_bytes(chunk))))
            chunk, size = [], 0
    if chunk:
        handles.append((chunk[-1][0], add_block(_block_bytes(chunk))))
    meta = add_block(_block_bytes([]))
    index = add_block(_block_bytes(handles, restart_interval=1))
    foot = meta + index
    out.extend(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
