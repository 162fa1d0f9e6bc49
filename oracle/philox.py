"""Counter-based noise stream (oracle; test infrastructure only).

The reference draws noise with ``tf.random.normal`` from TF's global stateful generator
(networks/dm3d.py:513,519) which cannot be reproduced outside TF.  The build therefore
DEFINES its production noise stream and this file is its CPU statement:

  Philox4x32-10 (Salmon et al., SC'11; Random123 known-answer vectors pinned in
  tests/test_philox.py), key = (seed_lo, seed_hi),
  counter = (element_index // 4, step, global_sample_index, stream)   stream 0 = per-step
  noise, stream 1 = x_T; the four 32-bit outputs feed two Box-Muller pairs:
      u1 = float(r >> 8) * 2^-24 + 2^-25   (0 < u1 <= 1)     u2 = float(r' >> 8) * 2^-24
      z0 = sqrt(-2 ln u1) cos(2 pi u2)     z1 = sqrt(-2 ln u1) sin(2 pi u2)
  element e of a sample takes z[e % 4] of counter e // 4.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10.  c* are uint32 arrays (broadcastable); returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    u32 = lambda a: np.asarray(a, dtype=np.uint32)  # noqa: E731
    return u32(c0), u32(c1), u32(c2), u32(c3)


def _box_muller(ra, rb):
    u1 = (ra >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24) + np.float32(2.0 ** -25)
    u2 = (rb >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1))
    ang = np.float64(2.0 * np.pi) * u2.astype(np.float64)
    return (rad * np.cos(ang).astype(np.float32)).astype(np.float32), \
           (rad * np.sin(ang).astype(np.float32)).astype(np.float32)


def normal(seed: int, step: int, sample_ids, n_elem: int, stream: int = 0) -> np.ndarray:
    """(len(sample_ids), n_elem) float32 standard normals of the production stream."""
    sample_ids = np.asarray(sample_ids, dtype=np.uint32)
    n4 = (n_elem + 3) // 4
    ctr = np.arange(n4, dtype=np.uint32)[None, :]
    sid = sample_ids[:, None]
    r0, r1, r2, r3 = philox4x32_10(ctr, np.uint32(step), sid, np.uint32(stream),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z0, z1 = _box_muller(r0, r1)
    z2, z3 = _box_muller(r2, r3)
    z = np.stack([z0, z1, z2, z3], axis=-1).reshape(len(sample_ids), n4 * 4)
    return np.ascontiguousarray(z[:, :n_elem])
