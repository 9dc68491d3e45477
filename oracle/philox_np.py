"""CPU oracle for the on-device collocation sampler -- numpy restatement of Philox4x32-10.

TEST INFRASTRUCTURE ONLY (see oracle/ref_port.py header): only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import this.

The reference draws its collocation points with torch.rand / rand_like (heat.py:125-126, simple_ode.py:91,
fitzhugh_nagumo.py:129, fredholm.py:67,100), i.e. with torch's Philox stream -- third-party arithmetic whose stream
layout is not part of the reference.  SURVEY 8f N2 asks for statistical, not bitwise, parity with it; what CAN be pinned
bit for bit is the generator itself: Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw: "Parallel random numbers:
as easy as 1, 2, 3", SC'11; Random123 `philox4x32_R(10, ctr, key)`), multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key
increments 0x9E3779B9 / 0xBB67AE85.  Parity pin: the three known-answer vectors of Random123's kat_vectors for
philox4x32-10 (tests/test_oracle.py::test_philox_kat); the CUDA kernel is compared with this file bit for bit.

Stream layout of dgmk_sample_uniform / dgmk_sample_heat (include/dgmk.h): element i of an output takes word i % 4 of
block counter (lo32(i / 4), hi32(i / 4), lo32(step), stream_id) under key (lo32(seed), hi32(seed));
u = (word >> 8) * 2^-24 in [0, 1); value = lo + (hi - lo) * u in FP32 (one FMA-free multiply-add: fl(fl((hi-lo)*u)+lo)).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: [..., 4] uint32, key: [..., 2] uint32 (broadcastable) -> [..., 4] uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32)
    k1 = np.asarray(key[..., 1], dtype=np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def words(n, seed, stream_id, step):
    """The first n 32-bit words of stream (seed, stream_id, step) in element order."""
    nb = (n + 3) // 4
    blk = np.arange(nb, dtype=np.uint64)
    ctr = np.empty((nb, 4), dtype=np.uint32)
    ctr[:, 0] = (blk & MASK).astype(np.uint32)
    ctr[:, 1] = (blk >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(int(step) & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32(int(stream_id) & 0xFFFFFFFF)
    key = np.array([int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key[None, :]).reshape(-1)[:n]


def uniform(n, lo, hi, seed, stream_id, step):
    """dgmk_sample_uniform: float32 [n]."""
    u = (words(n, seed, stream_id, step) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    span = np.float32(np.float32(hi) - np.float32(lo))
    return (span * u).astype(np.float32) + np.float32(lo)


def heat(B, xmax, tmax, xbd2, seed, step):
    """dgmk_sample_heat: x = xmax * u (stream 0), t = tmax * u' (stream 1); returns X, X0, XBD1, XBD2 as [B, 2] float32."""
    x = uniform(B, 0.0, xmax, seed, 0, step)
    t = uniform(B, 0.0, tmax, seed, 1, step)
    z = np.zeros(B, dtype=np.float32)
    return (np.stack([x, t], 1), np.stack([x, z], 1), np.stack([z, t], 1), np.stack([z + np.float32(xbd2), t], 1))
