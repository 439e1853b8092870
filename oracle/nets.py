"""CPU restatement of ac_nets.py in closed form (numpy).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows the reference:
  * NeuralNet.forward: in->H->H->out, ReLU, optional softmax .... ac_nets.py:26-41 (hidden_size=6, :24)
  * CriticNetwork.batch_update: one-hot select, MSE(target, Q[a]) .. ac_nets.py:62-72
  * ActorNetwork.batch_update: Categorical log-prob / entropy loss .. ac_nets.py:112-119
    (no zero_grad: gradients accumulate across updates)
  * torch.distributions.Categorical(probs=p) [third party, torch 2.0.1 pinned]:
    q = p / sum(p); logits = log(clamp(q, eps, 1-eps)), eps = 2^-23;
    log_prob = logits[a]; entropy = -sum(logits * q)
  * torch.optim.Adam single-tensor path [third party]: SURVEY.md Appendix A.5

No autograd is used here: gradients are the closed forms of SURVEY.md Appendix
A.4, so the oracle is independent of torch.  ``dtype=np.float32`` mimics the
reference's precision; ``np.float64`` gives a tighter ground truth for judging
the 1e-5 tolerance.

Pinned against: tests/golden/acnets_updates.npz (real classes: values, losses,
autograd gradients, post-Adam parameters, Q2 accumulation, Q8 target-with-grad)
and the update tapes in tests/golden/ia2c_*.npz, a2c_org.npz.
"""
from __future__ import annotations

import numpy as np

HIDDEN = 6
EPS_CLAMP = float(np.finfo(np.float32).eps)  # torch clamp_probs eps for float32


def n_params(F, O, H=HIDDEN):
    return H * F + H + H * H + H + O * H + O


def unpack(flat, F, O, H=HIDDEN):
    flat = np.asarray(flat)
    o = 0
    out = []
    for shape in ((H, F), (H,), (H, H), (H,), (O, H), (O,)):
        n = int(np.prod(shape))
        out.append(flat[o:o + n].reshape(shape))
        o += n
    assert o == flat.size, (o, flat.size)
    return out


def pack(parts):
    return np.concatenate([np.asarray(p).reshape(-1) for p in parts])


def forward(flat, x, F, O, softmax=False, H=HIDDEN, keep=False):
    W1, b1, W2, b2, W3, b3 = unpack(flat, F, O, H)
    x = np.asarray(x, dtype=flat.dtype).reshape(-1, F)
    z1 = x @ W1.T + b1
    h1 = np.maximum(z1, 0)
    z2 = h1 @ W2.T + b2
    h2 = np.maximum(z2, 0)
    y = h2 @ W3.T + b3
    out = y
    if softmax:
        e = np.exp(y - y.max(-1, keepdims=True))
        out = e / e.sum(-1, keepdims=True)
    if keep:
        return out, (x, z1, h1, z2, h2, y)
    return out


def backward(flat, cache, dy, F, O, H=HIDDEN):
    """Gradient of sum(dy * y) wrt the flat parameters (dy is wrt the pre-softmax output y)."""
    W1, b1, W2, b2, W3, b3 = unpack(flat, F, O, H)
    x, z1, h1, z2, h2, y = cache
    gW3 = dy.T @ h2
    gb3 = dy.sum(0)
    dh2 = dy @ W3
    dz2 = dh2 * (z2 > 0)
    gW2 = dz2.T @ h1
    gb2 = dz2.sum(0)
    dh1 = dz2 @ W2
    dz1 = dh1 * (z1 > 0)
    gW1 = dz1.T @ x
    gb1 = dz1.sum(0)
    return pack([gW1, gb1, gW2, gb2, gW3, gb3]).astype(flat.dtype)


def critic_loss_grad(flat, obs, act, target, F, O, next_obs=None, next_act=None, gamma_mask=None):
    """MSE(target, Q(obs)[act]) and its parameter gradient.

    If ``next_obs`` is given the target is ``target + gamma_mask * Q(next_obs)[next_act]`` WITH gradient
    through the second forward pass (ia2c.py:108-113, Q8; ``target`` then holds the rewards and
    ``gamma_mask`` = gamma (* mask)).  Returns (loss, grad, full_target).
    """
    dt = flat.dtype
    act = np.asarray(act).reshape(-1).astype(np.int64)
    target = np.asarray(target, dtype=dt).reshape(-1)
    Q, cache = forward(flat, obs, F, O, keep=True)
    B = Q.shape[0]
    rows = np.arange(B)
    if next_obs is not None:
        next_act = np.asarray(next_act).reshape(-1).astype(np.int64)
        gm = np.broadcast_to(np.asarray(gamma_mask, dtype=dt).reshape(-1), (B,)) if np.ndim(gamma_mask) else \
            np.full(B, gamma_mask, dtype=dt)
        Qn, cache_n = forward(flat, next_obs, F, O, keep=True)
        target = target + gm * Qn[rows, next_act]
    delta = target - Q[rows, act]
    loss = np.mean(delta * delta, dtype=dt)
    dQ = np.zeros_like(Q)
    dQ[rows, act] = -2.0 * delta / B
    grad = backward(flat, cache, dQ, F, O)
    if next_obs is not None:
        dQn = np.zeros_like(Qn)
        dQn[rows, next_act] = 2.0 * gm * delta / B
        grad = grad + backward(flat, cache_n, dQn, F, O)
    return loss.astype(dt), grad.astype(dt), target


def actor_loss_grad(flat, obs, act, adv, beta, F, O, adv_extra_dy=None):
    """mean(adv * (-log q[a]) - beta * H(q)) and its parameter gradient (adv treated as a constant).

    Returns (loss, grad, dloss_dadv) — the last is what autograd would send into ``adv``'s graph (Q7).
    """
    dt = flat.dtype
    act = np.asarray(act).reshape(-1).astype(np.int64)
    adv = np.asarray(adv, dtype=dt).reshape(-1)
    p, cache = forward(flat, obs, F, O, softmax=True, keep=True)
    B = p.shape[0]
    rows = np.arange(B)
    q = p / p.sum(-1, keepdims=True)
    eps = dt.type(EPS_CLAMP) if hasattr(dt, "type") else EPS_CLAMP
    inside = (q >= eps) & (q <= 1 - eps)
    logit = np.log(np.clip(q, eps, 1 - eps))
    neglogp = -logit[rows, act]
    ent = -(logit * q).sum(-1)
    loss = np.mean(adv * neglogp - beta * ent, dtype=dt)
    # dL/dq (per row, before the 1/B of the mean)
    g = beta * (logit + inside)
    g[rows, act] += -adv * inside[rows, act] / q[rows, act]
    dy = q * (g - (q * g).sum(-1, keepdims=True)) / B
    if adv_extra_dy is not None:
        dy = dy + adv_extra_dy
    grad = backward(flat, cache, dy.astype(dt), F, O)
    return loss.astype(dt), grad.astype(dt), (neglogp / B).astype(dt)


class AdamRef:
    """torch.optim.Adam defaults (betas .9/.999, eps 1e-8, no weight decay / amsgrad), Appendix A.5."""

    def __init__(self, n, lr, dtype=np.float32, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = np.zeros(n, dtype=dtype)
        self.v = np.zeros(n, dtype=dtype)
        self.t = 0

    def step(self, params, grad):
        dt = params.dtype.type
        self.t += 1
        g = grad.astype(params.dtype)
        self.m = self.m + (g - self.m) * dt(1 - self.b1)
        self.v = self.v * dt(self.b2) + dt(1 - self.b2) * g * g
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        step_size = self.lr / bc1
        denom = np.sqrt(self.v) / dt(bc2 ** 0.5) + dt(self.eps)
        return (params - dt(step_size) * (self.m / denom)).astype(params.dtype)


def sample_inverse_cdf(probs, u):
    """Performance-mode action sampler of the CUDA path (NOT the reference's torch.multinomial stream):
    q = p/sum(p) (Categorical renormalises); first k with u*sum(p) < cumsum(p)[k]; none -> last index.
    Sequential sums in the dtype of ``probs``, like the kernel."""
    p = np.asarray(probs)
    dt = p.dtype
    s = np.zeros(p.shape[:-1], dtype=dt)
    for k in range(p.shape[-1]):
        s = (s + p[..., k]).astype(dt)
    us = (np.asarray(u, dtype=dt).reshape(s.shape) * s).astype(dt)
    c = np.zeros_like(s)
    out = np.full(s.shape, p.shape[-1] - 1, dtype=np.int64)
    found = np.zeros(s.shape, dtype=bool)
    for k in range(p.shape[-1]):
        c = (c + p[..., k]).astype(dt)
        hit = (~found) & (us < c)
        out[hit] = k
        found |= hit
    return out
