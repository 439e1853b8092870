"""Belief filter on the GPU behind the reference's ``BeliefFilter`` class API.

Reference interface mirrored here (thinclab/IA2C):
  * ``BeliefFilter(num_models, num_actions, num_envs)`` with attributes ``prior`` [E,M],
    ``filterAction`` [M,A] (random row-stochastic), ``filters`` = its transpose, ``num_actions``,
    ``num_models``                                               belief_filter_deprecated.py:28-43
  * ``update(obs [E,A], prev_belief [E,M]) -> (ap int64[E], bprime [E,M], prediction [E,A])``
                                                                 belief_filter_deprecated.py:45-59
The constructor draws its models with ``np.random.rand`` and ``update`` draws ``u`` with
``np.random.rand(E, 1)`` exactly where the reference does, so a script run under the same numpy seed
consumes the host RNG stream identically (this is what makes replay parity possible).  The arithmetic
is the fp64 kernel ``ia2c_belief_update_dense`` (csrc/belief.cu); there is no CPU implementation here.
"""
from __future__ import annotations

import numpy as np

from . import _lib

# Known models quoted in the reference's comments (belief_filter_deprecated.py:31-37)
ORG_KNOWN_FILTER_ACTION = np.array([[0.8, 0.1, 0.1], [0.6, 0.2, 0.2], [0.4, 0.3, 0.3]])
HVT_KNOWN_FILTER_ACTION = np.array([[0.8, 0.05, 0.05, 0.05, 0.05], [0.6, 0.1, 0.1, 0.1, 0.1],
                                    [0.4, 0.15, 0.15, 0.15, 0.15], [0.2, 0.2, 0.2, 0.2, 0.2],
                                    [0.1, 0.225, 0.225, 0.225, 0.225]])


def generate_random_probability_matrix(m, n):
    """belief_filter_deprecated.py:22-25 (constructor-time host code, same numpy calls)."""
    matrix = np.random.rand(m, n)
    matrix /= np.sum(matrix, axis=1)[:, np.newaxis]
    return matrix


class BeliefFilter:
    def __init__(self, num_models, num_actions, num_envs, known=None):
        torch = _lib.require_cuda()
        self._torch = torch
        self.lib = _lib.load()
        self.prior = np.tile(np.ones(num_models) * round(1.0 / num_models, 2), (num_envs, 1))
        if known is None:
            self.filterAction = generate_random_probability_matrix(num_models, num_actions)
        else:  # f3: the papers' known models ("org" / "hvt") or an explicit [M,A] matrix
            fa = {"org": ORG_KNOWN_FILTER_ACTION, "hvt": HVT_KNOWN_FILTER_ACTION}.get(known, known)
            self.filterAction = np.array(fa, dtype=np.float64)
            assert self.filterAction.shape == (num_models, num_actions)
        self.filters = self.filterAction.transpose()
        self.num_actions = num_actions
        self.num_models = num_models
        self._dev = torch.device(f"cuda:{torch.cuda.current_device()}")
        self._cap = 0

    def _buffers(self, E):
        torch = self._torch
        M, A = self.num_models, self.num_actions
        if E > self._cap:
            n_in = M * A + E * A + E * M + E
            n_out = E * M + E * A + E
            self._h_in = torch.empty(n_in, dtype=torch.float64).pin_memory()
            self._d_in = torch.empty(n_in, dtype=torch.float64, device=self._dev)
            self._d_out = torch.empty(n_out, dtype=torch.float64, device=self._dev)
            self._h_out = torch.empty(n_out, dtype=torch.float64).pin_memory()
            self._cap = E
        return self._h_in, self._d_in, self._d_out, self._h_out

    def update(self, obs, prev_belief):
        torch = self._torch
        obs = np.asarray(obs, dtype=np.float64)
        prev = np.asarray(prev_belief, dtype=np.float64)
        E = obs.shape[0]
        M, A = self.num_models, self.num_actions
        u = np.random.rand(E, 1)  # same host draw as belief_filter_deprecated.py:56
        h_in, d_in, d_out, h_out = self._buffers(E)
        hin = h_in.numpy()
        o_fa, o_lik, o_prev, o_u = 0, M * A, M * A + E * A, M * A + E * A + E * M
        hin[o_fa:o_lik] = np.asarray(self.filterAction, dtype=np.float64).reshape(-1)  # may have been replaced
        hin[o_lik:o_prev] = obs.reshape(-1)
        hin[o_prev:o_u] = prev.reshape(-1)
        hin[o_u:o_u + E] = u.reshape(-1)
        n_in = o_u + E
        with torch.cuda.device(self._dev):
            d_in[:n_in].copy_(h_in[:n_in], non_blocking=True)
            base_in, base_out = d_in.data_ptr(), d_out.data_ptr()
            p_b, p_pred, p_ap = base_out, base_out + 8 * E * M, base_out + 8 * (E * M + E * A)
            _lib.check(self.lib.ia2c_belief_update_dense(
                base_in + 8 * o_fa, base_in + 8 * o_lik, base_in + 8 * o_prev, base_in + 8 * o_u,
                p_ap, p_b, p_pred, E, M, A, _lib.stream_ptr()), "ia2c_belief_update_dense")
            n_out = E * M + E * A + E
            h_out[:n_out].copy_(d_out[:n_out], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        hout = h_out.numpy()
        bprime = hout[:E * M].reshape(E, M).copy()
        prediction = hout[E * M:E * M + E * A].reshape(E, A).copy()
        ap = hout[E * M + E * A:n_out].view(np.int64).copy()
        return ap, bprime, prediction
