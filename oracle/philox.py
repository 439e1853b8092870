"""Philox4x32-10 counter-based generator in numpy.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Not part of the reference (which uses torch.multinomial / np.random.rand streams that cannot be
reproduced on a GPU).  It mirrors the generator inside the CUDA rollout kernels
(ia2c_b200/csrc/common.cuh) so that performance-mode rollouts can be replayed bit-exactly:
the oracle regenerates the uniforms the kernel drew and injects them as a tape.

Known-answer pin: Random123's published vectors for philox4x32-10
(counter=0,key=0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8; all-ones -> 408f276d 41c83b0e a20bc7c6 6d5451fd).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)

STREAM_ACTION = 1
STREAM_BELIEF = 2


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def draw(seed, stream, episode, t, index):
    """The counter layout used by the CUDA kernels: key = seed (lo, hi); counter =
    (index lo, index hi, t | stream << 16, episode)."""
    index = np.asarray(index, dtype=np.uint64)
    c0 = (index & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c1 = (index >> np.uint64(32)).astype(np.uint32)
    c2 = np.uint32((int(t) & 0xFFFF) | (int(stream) << 16))
    c3 = np.uint32(int(episode) & 0xFFFFFFFF)
    return philox4x32(c0, c1, c2, c3, int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)


def uniform_f32(seed, stream, episode, t, index):
    """24-bit uniform in [0,1) as float32 (action sampler)."""
    x0, _, _, _ = draw(seed, stream, episode, t, index)
    return ((x0 >> np.uint32(8)).astype(np.float32)) * np.float32(2.0 ** -24)


def _unit_f64(hi, lo):
    bits = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return (bits >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def uniform_f64(seed, stream, episode, t, index):
    """53-bit uniform in [0,1) as float64 from words (x0, x1) (same resolution as np.random.rand)."""
    x0, x1, _, _ = draw(seed, stream, episode, t, index)
    return _unit_f64(x0, x1)


def belief_uniforms(seed, episode, t, rows, K):
    """The belief sampler's uniforms u[..., jj] for belief rows `rows` (= env * N + agent, any shape) and K modelled
    others.  One Philox block serves FOUR slots (common.cuh: philox_belief_quad): slots 4s .. 4s+3 of row r share the
    draw at index r * ceil(K/4) + s, word w -> slot 4s + w, and a slot's uniform is the centred 32-bit value
    (word + 0.5) * 2^-32 (exact in float64; common.cuh: belief_word_to_unit_f64)."""
    rows = np.asarray(rows, dtype=np.uint64)
    kq = (K + 3) // 4
    idx = rows[..., None] * np.uint64(kq) + np.arange(kq, dtype=np.uint64)
    words = np.stack(draw(seed, STREAM_BELIEF, episode, t, idx), axis=-1).reshape(rows.shape + (4 * kq,))
    return ((words.astype(np.float64) + 0.5) * 2.0 ** -32)[..., :K]
