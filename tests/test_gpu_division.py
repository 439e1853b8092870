"""The library's branch-free fp64 division sequence (csrc/common.cuh: ddiv_seq) must equal IEEE division bit for
bit on the operand ranges the belief filter (bp/S, k/100) and the Org reward recurrence (r/10) produce."""
import numpy as np
import pytest

from tests.helpers import dev, host

pytestmark = pytest.mark.gpu


def _divide(a, b):
    import torch
    from ia2c_b200 import _lib
    lib = _lib.load()
    ta, tb = dev(a), dev(b)
    qs, qi = torch.empty_like(ta), torch.empty_like(ta)
    _lib.check(lib.ia2c_debug_divide(_lib.ptr(ta), _lib.ptr(tb), _lib.ptr(qs), _lib.ptr(qi), ta.numel(), _lib.stream_ptr()))
    return host(qs), host(qi)


def test_division_sequence_is_ieee_exact():
    rng = np.random.RandomState(0)
    n = 4_000_000
    cases = []
    # belief: numerators = sums of products of likelihood x model prob x prior (incl. exact zeros), denominators their sums
    fa = rng.rand(n, 3)
    fa /= fa.sum(1, keepdims=True)
    prior = rng.randint(0, 101, n) / 100.0
    lik = np.where(rng.rand(n, 3) < 0.34, 0.8, 0.1)
    bp = (lik[:, 0] * (fa[:, 0] * prior) + lik[:, 1] * (fa[:, 1] * prior)) + lik[:, 2] * (fa[:, 2] * prior)
    S = bp + rng.rand(n) * 0.9 + 1e-3
    cases.append((bp, S))
    cases.append((rng.rand(n), rng.rand(n) * 1.1 + 1e-3))
    # rounding: k / 100 for every k the filter can produce (and random integers)
    cases.append((np.arange(0, 101, dtype=np.float64), np.full(101, 100.0)))
    cases.append((rng.randint(0, 10 ** 6, n).astype(np.float64), np.full(n, 100.0)))
    # reward recurrence: r / 10 for r in the reachable range, random mantissas, and exact zero
    # (a = -0.0 is outside the domain: the sequence returns +0.0 where IEEE gives -0.0; r starts at +0.0 and
    #  base + r/10 cannot produce -0.0 in round-to-nearest, and belief numerators are products of non-negatives)
    cases.append((np.concatenate([[0.0], rng.uniform(-112, 7, n)]), np.full(n + 1, 10.0)))
    r = np.zeros(200000)
    vals = [r.copy()]
    for _ in range(40):  # actual trajectories of r <- base + r/10
        base = rng.choice([-100.0, 1.0, 5.0, 6.0], size=r.size)
        r = base + r / 10.0
        vals.append(r.copy())
    cases.append((np.concatenate(vals), np.full(200000 * 41, 10.0)))
    for a, b in cases:
        qs, qi = _divide(a, b)
        assert np.array_equal(qi, a / b)                   # the device's IEEE division equals the host's
        assert np.array_equal(qs.view(np.int64), qi.view(np.int64)), int((qs != qi).sum())
