"""CPU restatement of the belief filter.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows ``BeliefFilter`` in the reference:
  * prior = round(1/M, 2) per model ............. belief_filter_deprecated.py:29
  * models: row-stochastic filterAction [M,A] ... belief_filter_deprecated.py:22-25,39-40
  * update ...................................... belief_filter_deprecated.py:45-59
  * likelihood 0.8 on the observed action, 0.1 elsewhere (no RNG) ... ia2c.py:53-58

The update is written as the explicit fp64 operation sequence the CUDA kernel
implements: one rounding per product, left-to-right
sums, IEEE division, round-half-even to 2 decimals.  ``prediction`` is produced
by a BLAS matmul in the reference, so it is only pinned to a few ulp; ``ap`` and
``bprime`` are pinned bit-exactly.

Pinned against: tests/golden/belief_vectors.npz (real class, M in {3,5}, A in {3,5},
random + the papers' known models) and the bf tapes in tests/golden/ia2c_*.npz.
"""
from __future__ import annotations

import numpy as np

LIK_HIT = 0.8
LIK_MISS = 0.1


def uniform_prior(num_envs, num_models):
    return np.tile(np.ones(num_models) * round(1.0 / num_models, 2), (num_envs, 1))


def likelihood_from_action(other_action, num_actions=3):
    """ia2c.py:53-58 — deterministic 'noisy' private observation."""
    other_action = np.asarray(other_action)
    lik = np.full(other_action.shape + (num_actions,), LIK_MISS, dtype=np.float64)
    np.put_along_axis(lik, other_action[..., None].astype(np.int64), LIK_HIT, axis=-1)
    return lik


def belief_update(filter_action, lik, prev, u):
    """filter_action f64[M,A]; lik f64[R,A]; prev f64[R,M]; u f64[R] -> (ap i64[R], bprime[R,M], prediction[R,A]).

    R is any number of rows (envs, or env x agent x modelled-other pairs).
    """
    fa = np.asarray(filter_action, dtype=np.float64)
    lik = np.asarray(lik, dtype=np.float64)
    prev = np.asarray(prev, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64).reshape(-1)
    M, A = fa.shape
    bp = None
    for a in range(A):  # left-to-right over actions, each product rounded once
        term = lik[:, a:a + 1] * (fa[:, a][None, :] * prev)
        bp = term if bp is None else bp + term
    S = bp[:, 0].copy()
    for m in range(1, M):
        S = S + bp[:, m]
    b = bp / S[:, None]
    pred = np.zeros((lik.shape[0], A))
    for m in range(M):
        pred = pred + b[:, m:m + 1] * fa[m][None, :]
    c = pred[:, 0].copy()
    ap = np.zeros(lik.shape[0], dtype=np.int64)
    found = u < c
    for a in range(1, A):  # first a with u < cumsum; none -> 0 (Q11)
        c = c + pred[:, a]
        hit = (~found) & (u < c)
        ap[hit] = a
        found |= hit
    bprime = np.rint(b * 100.0) / 100.0  # numpy .round(2): rint(x*100)/100 (Q10)
    return ap, bprime, pred


def to_hundredths(b):
    """Rounded posteriors are k/100 with integer k (lossless uint8 storage)."""
    k = np.rint(np.asarray(b) * 100.0).astype(np.int64)
    assert np.array_equal(k / 100.0, np.asarray(b)), "not a 2-decimal belief"
    return k.astype(np.uint8)


def from_hundredths(k):
    return np.asarray(k, dtype=np.float64) / 100.0
