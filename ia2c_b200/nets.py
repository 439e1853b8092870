"""Actor / critic networks on the GPU behind the reference's ``ac_nets`` class API.

Reference interface mirrored here (thinclab/IA2C, ac_nets.py):
  * ``hidden_size = 6``                                                         :24
  * ``NeuralNet(state_dim, action_dim, b_actor=False)`` with ``l1, l2, l3``      :26-41
  * ``CriticNetwork(name, n_features, critic_actions, lr, cuda=False)``:
    ``net, loss, optimizer, num_outs, cuda, losses, critic_loss``;
    ``run_main(obs, grad=False)``, ``batch_update(obs, act, target, action_distribution=False)``  :43-80
  * ``ActorNetwork(name, n_features, actor_actions, lr, beta, cuda=False)``:
    ``net, optimizer, num_outs, cuda, beta, losses, actor_loss``;
    ``sample_action(obs, grad=False)``, ``action_distribution(obs, grad=False)``,
    ``batch_update(obs, act, adv, retain=False)`` (never zeroes its gradients, Q2)            :83-127

Every forward, backward, loss, sampler and optimiser step is a kernel of libia2c_b200.so.  torch is
used for device memory, and its autograd ENGINE only as plumbing between our custom Functions and the
ad-hoc graphs the reference's scripts build around them (targets / advantages that carry gradient:
SURVEY.md Q7, Q8).  Inputs may live on the CPU (the scripts create CPU tensors); outputs are returned
on the input's device.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib

hidden_size = 6


def _check_hidden_size():
    """The kernels are compiled for the reference's 6-wide hidden layers (IA2C_HIDDEN, ac_nets.py:24).  A script that
    rebinds ``hidden_size`` — here or on the ``ac_nets`` drop-in module it star-imports from — would be silently ignored;
    refuse instead."""
    import sys

    mods = [sys.modules[__name__]] + [m for n, m in list(sys.modules.items()) if n == "ac_nets" and m is not None]
    for m in mods:
        v = getattr(m, "hidden_size", 6)
        if v != 6:
            raise _lib.IA2CError(f"{m.__name__}.hidden_size = {v!r}: libia2c_b200.so is compiled for hidden_size = 6 "
                                 "(IA2C_HIDDEN in include/ia2c_b200.h); other widths are not supported")


def n_params(state_dim, action_dim):
    return hidden_size * state_dim + hidden_size + hidden_size * hidden_size + hidden_size + action_dim * hidden_size + action_dim


def _device():
    _lib.require_cuda()
    return torch.device(f"cuda:{torch.cuda.current_device()}")


class _MLPFunction(torch.autograd.Function):
    """y = NeuralNet(x) on device tensors; backward = ia2c_mlp_backward (closed form, no torch ops)."""

    @staticmethod
    def forward(ctx, x, flat, state_dim, action_dim, softmax):
        lib = _lib.load()
        x2 = x.detach().reshape(-1, state_dim).to(torch.float32).contiguous()
        rows = x2.shape[0]
        y = torch.empty(rows, action_dim, dtype=torch.float32, device=x2.device)
        # keep the layer-1 activations when a backward pass may follow: it then never re-reads x to recompute them
        h1 = torch.empty(rows, hidden_size, dtype=torch.float32, device=x2.device) if any(ctx.needs_input_grad) else None
        _lib.check(lib.ia2c_mlp_forward(_lib.ptr(flat.detach()), _lib.ptr(x2), _lib.ptr(y), _lib.ptr(h1), rows, state_dim,
                                        action_dim, 1, int(softmax), _lib.stream_ptr()), "ia2c_mlp_forward")
        ctx.save_for_backward(x2, flat, h1)
        ctx.dims = (state_dim, action_dim, softmax, tuple(x.shape))
        return y.reshape(*x.shape[:-1], action_dim)

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x2, flat, h1 = ctx.saved_tensors
        state_dim, action_dim, softmax, xshape = ctx.dims
        rows = x2.shape[0]
        dy = gy.reshape(rows, action_dim).to(torch.float32).contiguous()
        grad = torch.empty_like(flat)
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        ws = torch.empty(lib.ia2c_mlp_backward_workspace(rows, state_dim, action_dim), dtype=torch.float32,
                         device=x2.device)
        _lib.check(lib.ia2c_mlp_backward(_lib.ptr(flat.detach()), _lib.ptr(x2), _lib.ptr(dy), _lib.ptr(h1), _lib.ptr(grad),
                                         _lib.ptr(dx), _lib.ptr(ws), rows, state_dim, action_dim, int(softmax), 0,
                                         _lib.stream_ptr()), "ia2c_mlp_backward")
        return (dx.reshape(xshape) if dx is not None else None), grad, None, None, None


def _is_index_input(s):
    """Integer-typed observations are class indices standing for one_hot(idx, state_dim) rows (the
    a2c_test.py shape, where the reference materialises F.one_hot(states, 500).float()); floats are dense features."""
    if isinstance(s, torch.Tensor):
        return not (s.dtype.is_floating_point or s.dtype.is_complex or s.dtype == torch.bool)
    return np.asarray(s).dtype.kind in "iu"


def _checked_indices(idx, state_dim, dev):
    i = torch.as_tensor(idx).to(dev, torch.int64)
    flat = i.reshape(-1).contiguous()
    lo, hi = torch.stack(torch.aminmax(flat)).tolist() if flat.numel() else (0, 0)   # one reduction, one sync
    if lo < 0 or hi >= state_dim:   # F.one_hot raises here too
        raise RuntimeError(f"Class values must be in [0, {state_dim}) for an index-typed observation")
    return i, flat


class _MLPIndexFunction(torch.autograd.Function):
    """y = NeuralNet(one_hot(idx)) without the one-hot tensor: ia2c_mlp_forward_index / ia2c_mlp_backward_index
    (bit-identical to _MLPFunction on the materialised rows)."""

    @staticmethod
    def forward(ctx, idx_flat, flat, state_dim, action_dim, softmax):
        lib = _lib.load()
        rows = idx_flat.shape[0]
        y = torch.empty(rows, action_dim, dtype=torch.float32, device=flat.device)
        h1 = torch.empty(rows, hidden_size, dtype=torch.float32, device=flat.device) if any(ctx.needs_input_grad) else None
        _lib.check(lib.ia2c_mlp_forward_index(_lib.ptr(flat.detach()), _lib.ptr(idx_flat), _lib.ptr(y), _lib.ptr(h1), rows,
                                              state_dim, action_dim, int(softmax), _lib.stream_ptr()), "ia2c_mlp_forward_index")
        ctx.save_for_backward(idx_flat, flat, h1)
        ctx.dims = (state_dim, action_dim, softmax)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        idx_flat, flat, h1 = ctx.saved_tensors
        state_dim, action_dim, softmax = ctx.dims
        rows = idx_flat.shape[0]
        dy = gy.reshape(rows, action_dim).to(torch.float32).contiguous()
        grad = torch.empty_like(flat)
        ws = torch.empty(lib.ia2c_mlp_backward_workspace(rows, state_dim, action_dim), dtype=torch.float32, device=flat.device)
        _lib.check(lib.ia2c_mlp_backward_index(_lib.ptr(flat.detach()), _lib.ptr(idx_flat), _lib.ptr(dy), _lib.ptr(h1),
                                               _lib.ptr(grad), _lib.ptr(ws), rows, state_dim, action_dim, int(softmax), 0,
                                               _lib.stream_ptr()), "ia2c_mlp_backward_index")
        return None, grad, None, None, None


class _CriticLossFunction(torch.autograd.Function):
    """mean((target - Q[act])^2): ia2c_critic_loss gives the loss, dQ and dtarget in one kernel."""

    @staticmethod
    def forward(ctx, Q, act, target):
        lib = _lib.load()
        O = Q.shape[-1]
        q2 = Q.detach().reshape(-1, O).contiguous()
        B = q2.shape[0]
        a = act.reshape(-1).to(torch.int32).contiguous()
        t = target.detach().reshape(-1).to(torch.float32).contiguous()
        if a.numel() != B or t.numel() != B:
            raise ValueError(f"critic loss: {B} rows of Q but {a.numel()} actions / {t.numel()} targets")
        loss = torch.empty((), dtype=torch.float32, device=q2.device)
        dQ = torch.empty_like(q2)
        dT = torch.empty_like(t)
        ws = torch.empty(lib.ia2c_loss_workspace(B), dtype=torch.float32, device=q2.device)
        _lib.check(lib.ia2c_critic_loss(_lib.ptr(q2), _lib.ptr(a), _lib.ptr(t), _lib.ptr(loss), _lib.ptr(dQ),
                                        _lib.ptr(dT), _lib.ptr(ws), B, O, _lib.stream_ptr()), "ia2c_critic_loss")
        ctx.save_for_backward(dQ, dT)
        ctx.shapes = (tuple(Q.shape), tuple(target.shape))
        return loss

    @staticmethod
    def backward(ctx, g):
        dQ, dT = ctx.saved_tensors
        qs, ts = ctx.shapes
        return (dQ * g).reshape(qs), None, (dT * g).reshape(ts) if ctx.needs_input_grad[2] else None


class _ActorLossFunction(torch.autograd.Function):
    """mean(adv * (-log q[a]) - beta * H(q)) with Categorical(probs=p) semantics: ia2c_actor_loss."""

    @staticmethod
    def forward(ctx, probs, act, adv, beta):
        lib = _lib.load()
        O = probs.shape[-1]
        p2 = probs.detach().reshape(-1, O).contiguous()
        B = p2.shape[0]
        a = act.reshape(-1).to(torch.int32).contiguous()
        ad = adv.detach().reshape(-1).to(torch.float32).contiguous()
        if a.numel() != B or ad.numel() != B:
            raise ValueError(f"actor loss: {B} rows of probs but {a.numel()} actions / {ad.numel()} advantages")
        loss = torch.empty((), dtype=torch.float32, device=p2.device)
        dP = torch.empty_like(p2)
        dA = torch.empty_like(ad)
        status = torch.zeros(1, dtype=torch.int32, device=p2.device)
        ws = torch.empty(lib.ia2c_loss_workspace(B), dtype=torch.float32, device=p2.device)
        _lib.check(lib.ia2c_actor_loss(_lib.ptr(p2), _lib.ptr(a), _lib.ptr(ad), float(beta), _lib.ptr(loss),
                                       _lib.ptr(dP), _lib.ptr(dA), _lib.ptr(status), _lib.ptr(ws), B, O,
                                       _lib.stream_ptr()), "ia2c_actor_loss")
        ctx.save_for_backward(dP, dA)
        ctx.shapes = (tuple(probs.shape), tuple(adv.shape))
        _ActorLossFunction.last_status = status
        return loss

    @staticmethod
    def backward(ctx, g):
        dP, dA = ctx.saved_tensors
        ps, as_ = ctx.shapes
        return (dP * g).reshape(ps), None, (dA * g).reshape(as_) if ctx.needs_input_grad[2] else None, None


class _LinearView:
    """``net.l1`` / ``l2`` / ``l3``: weight and bias views into the flat parameter vector."""

    def __init__(self, flat, w_off, out_f, in_f):
        self.in_features, self.out_features = in_f, out_f
        self._flat, self._w, self._b = flat, w_off, w_off + out_f * in_f

    @property
    def weight(self):
        return self._flat.data[self._w:self._b].view(self.out_features, self.in_features)

    @property
    def bias(self):
        return self._flat.data[self._b:self._b + self.out_features]


class NeuralNet(nn.Module):
    """in -> 6 -> 6 -> out MLP (ReLU, optional softmax) whose parameters are ONE flat CUDA vector in
    nn.Linear state_dict order.  Initial values are drawn exactly as the reference draws them
    (three nn.Linear default inits from torch's CPU generator, in l1, l2, l3 order)."""

    def __init__(self, state_dim, action_dim, b_actor=False):
        super().__init__()
        _check_hidden_size()
        dev = _device()
        self.state_dim, self.action_dim, self.b_actor = int(state_dim), int(action_dim), bool(b_actor)
        l1 = nn.Linear(state_dim, hidden_size)
        l2 = nn.Linear(hidden_size, hidden_size)
        l3 = nn.Linear(hidden_size, action_dim)
        flat = torch.cat([t.detach().reshape(-1) for l in (l1, l2, l3) for t in (l.weight, l.bias)])
        self.flat = nn.Parameter(flat.to(dev))
        o1 = 0
        o2 = o1 + hidden_size * state_dim + hidden_size
        o3 = o2 + hidden_size * hidden_size + hidden_size
        self.l1 = _LinearView(self.flat, o1, hidden_size, state_dim)
        self.l2 = _LinearView(self.flat, o2, hidden_size, hidden_size)
        self.l3 = _LinearView(self.flat, o3, action_dim, hidden_size)

    def forward(self, s):
        """Float input: dense features [..., state_dim] (the reference's call).  Integer input: class indices [...]
        standing for one_hot(idx, state_dim) rows — same output bits, no [rows, state_dim] tensor (index fast path)."""
        dev = self.flat.device
        if _is_index_input(s):
            idx, flat_idx = _checked_indices(s, self.state_dim, dev)
            y = _MLPIndexFunction.apply(flat_idx, self.flat, self.state_dim, self.action_dim, self.b_actor)
            return y.reshape(*idx.shape, self.action_dim)
        x = s if (isinstance(s, torch.Tensor) and s.device == dev) else torch.as_tensor(s).to(dev)
        y = _MLPFunction.apply(x.float(), self.flat, self.state_dim, self.action_dim, self.b_actor)
        return y

    # state_dict in the reference's key layout
    def state_dict(self, *args, **kwargs):
        return {f"l{i}.{k}": getattr(getattr(self, f"l{i}"), k).detach().clone()
                for i in (1, 2, 3) for k in ("weight", "bias")}

    def load_state_dict(self, sd, strict=True):
        with torch.no_grad():
            for i in (1, 2, 3):
                for k in ("weight", "bias"):
                    getattr(getattr(self, f"l{i}"), k).copy_(torch.as_tensor(sd[f"l{i}.{k}"]))

    def load_flat(self, flat):
        with torch.no_grad():
            self.flat.copy_(torch.as_tensor(np.asarray(flat), dtype=torch.float32))


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam-compatible front (defaults betas .9/.999, eps 1e-8) whose ``step`` is the
    ``ia2c_adam_step`` kernel over flat parameter vectors.  The step counter lives on the device."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        lib = _lib.load()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise _lib.IA2CError("ia2c_b200.Adam only steps CUDA parameters")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p).reshape(-1)
                    st["exp_avg_sq"] = torch.zeros_like(p).reshape(-1)
                    st["step"] = torch.zeros(1, dtype=torch.int32, device=p.device)
                g = p.grad.reshape(-1).contiguous()
                _lib.check(lib.ia2c_adam_step(_lib.ptr(p.data.view(-1)), _lib.ptr(g), None, _lib.ptr(st["exp_avg"]),
                                              _lib.ptr(st["exp_avg_sq"]), _lib.ptr(st["step"]), group["lr"],
                                              group["betas"][0], group["betas"][1], group["eps"], 1, p.numel(),
                                              _lib.stream_ptr()), "ia2c_adam_step")


def _index_like(act, lead_shape, who):
    """The reference does ``act.squeeze(-1)``; with one env and a [T,E] action tensor that also squeezes the
    ENV axis and silently trains on a broadcast [T,T] loss.  We reject that case."""
    idx = act.squeeze(-1)
    if tuple(idx.shape) != tuple(lead_shape):
        raise ValueError(f"{who}: action tensor of shape {tuple(act.shape)} does not index outputs of shape "
                         f"{tuple(lead_shape)} after squeeze(-1) (with n_envs == 1 the reference's squeeze drops the "
                         f"env axis and trains on a broadcast [T,T] loss; this implementation refuses that input)")
    return idx


def _raise_on_status(status, state_dim):
    """status word of ia2c_net_update: bit 1 = an index-typed observation outside [0, state_dim) (F.one_hot raises for it too;
    here the update has already been applied with the offending rows clamped), bit 0 = Categorical's simplex validation."""
    if status & 2:
        raise RuntimeError(f"Class values must be in [0, {state_dim}) for an index-typed observation")
    if status & 1:  # the reference's Categorical validation raises here
        raise ValueError("Expected parameter probs of distribution Categorical to satisfy the constraint Simplex()")


def _no_graph(t):
    return not (isinstance(t, torch.Tensor) and t.requires_grad)


def _fused_update(net, optimizer, kind, obs, act, signal, beta, who):
    """batch_update without a framework graph (target / advantage carries none): ONE C call, ``ia2c_net_update`` —
    forward, loss, backward, Adam — on a cached workspace.  Same arithmetic as the autograd path (the same kernels, or
    for wide dense inputs the single-pass kernel that reads the observations once).  Shares the optimizer's state, so
    fused and autograd updates can alternate.  -> (loss as numpy float32, status word), read back in one pinned 8-byte copy, or
    None if not applicable."""
    group = optimizer.param_groups[0]
    if len(optimizer.param_groups) != 1 or tuple(group["betas"]) != (0.9, 0.999) or group["eps"] != 1e-8:
        return None
    lib = _lib.load()
    p = net.flat
    dev = p.device
    F_, O = net.state_dim, net.action_dim
    if _is_index_input(obs):   # the range check rides on the update (status bit 1): no extra reduction + host sync here
        idx = torch.as_tensor(obs).to(dev, torch.int64)
        flat_idx = idx.reshape(-1).contiguous()
        lead, x2, rows = tuple(idx.shape), None, flat_idx.numel()
    else:
        x = obs if (isinstance(obs, torch.Tensor) and obs.device == dev) else torch.as_tensor(obs).to(dev)
        if x.shape[-1] != F_:
            return None
        lead, flat_idx = tuple(x.shape[:-1]), None
        x2 = x.detach().reshape(-1, F_).to(torch.float32).contiguous()
        rows = x2.shape[0]
    a = _index_like(torch.as_tensor(act), lead, who).to(dev).reshape(-1).to(torch.int32).contiguous()
    sig = torch.as_tensor(signal).detach().to(dev, torch.float32).reshape(-1).contiguous()
    if rows == 0 or a.numel() != rows or sig.numel() != rows:
        return None
    st = optimizer.state[p]
    if not st:
        st["exp_avg"] = torch.zeros_like(p).reshape(-1)
        st["exp_avg_sq"] = torch.zeros_like(p).reshape(-1)
        st["step"] = torch.zeros(1, dtype=torch.int32, device=dev)
    if p.grad is None:
        p.grad = torch.zeros_like(p)          # the actor's running gradient sum starts at zero (Q2); the critic's is overwritten
    cache = net.__dict__.setdefault("_fused_cache", {})
    key = (rows, F_, O)
    if key not in cache:
        cache.clear()
        n = int(lib.ia2c_net_update_workspace(rows, F_, O))
        out = torch.zeros(2, dtype=torch.int32, device=dev)          # [loss (float bits), status]: one 8-byte read-back
        cache[key] = (torch.empty(n, dtype=torch.float32, device=dev), out[:1].view(torch.float32), out[1:], out,
                      torch.zeros(2, dtype=torch.int32).pin_memory())
    ws, loss, status, out, h_out = cache[key]
    if kind == 1 or flat_idx is not None:
        status.zero_()
    with torch.no_grad():
        _lib.check(lib.ia2c_net_update(kind, _lib.ptr(p.data.view(-1)), _lib.ptr(p.grad.view(-1)), _lib.ptr(st["exp_avg"]),
                                       _lib.ptr(st["exp_avg_sq"]), _lib.ptr(st["step"]), _lib.ptr(x2), _lib.ptr(flat_idx), _lib.ptr(a),
                                       _lib.ptr(sig), float(beta), float(group["lr"]), _lib.ptr(loss), _lib.ptr(status), _lib.ptr(ws),
                                       rows, F_, O, _lib.stream_ptr()), "ia2c_net_update")
    h_out.copy_(out, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    host = h_out.numpy()
    return host[:1].view(np.float32)[0].copy(), int(host[1]) if (kind == 1 or flat_idx is not None) else 0


class CriticNetwork:
    def __init__(self, name, n_features, critic_actions, lr, cuda=False):
        self.name = name
        self.num_outs = critic_actions
        self.net = NeuralNet(n_features, critic_actions)
        self.loss = nn.MSELoss()
        self.optimizer = Adam(self.net.parameters(), lr=lr)
        self.cuda = cuda
        self.losses = []

    def run_main(self, obs, grad=False):
        in_dev = obs.device if isinstance(obs, torch.Tensor) else torch.device("cpu")
        if not grad:
            with torch.no_grad():
                out = self.net(obs)
        else:
            out = self.net(obs)
        return out.to(in_dev)

    def batch_update(self, obs, act, target, action_distribution=False):
        dev = self.net.flat.device
        if not action_distribution and _no_graph(target) and _no_graph(obs):   # no graph to feed: one fused C call
            fused = _fused_update(self.net, self.optimizer, 0, obs, act, target, 0.0, "CriticNetwork.batch_update")
            if fused is not None:
                _raise_on_status(fused[1], self.net.state_dim)
                self.losses.append(fused[0])
                if len(self.losses) > 20:
                    del self.losses[0]
                self.critic_loss = np.mean(self.losses)
                return
        self.optimizer.zero_grad()
        Q = self.net.forward(obs)
        target_d = target.to(dev)  # differentiable copy: gradient flows back into the caller's graph (Q8)
        if not action_distribution:
            idx = _index_like(torch.as_tensor(act), Q.shape[:-1], "CriticNetwork.batch_update").to(dev)
            loss = _CriticLossFunction.apply(Q, idx, target_d)
        else:
            dot_prd = (Q * torch.as_tensor(act).to(dev)).sum(-1, keepdims=True)
            loss = _CriticLossFunction.apply(dot_prd, torch.zeros(dot_prd.shape[:-1], dtype=torch.int32, device=dev),
                                             target_d)
        loss.backward()
        self.optimizer.step()
        get_loss = loss.detach().cpu().numpy()
        self.losses.append(get_loss)
        if len(self.losses) > 20:
            del self.losses[0]
        self.critic_loss = np.mean(self.losses)


class ActorNetwork:
    _instances = 0

    def __init__(self, name, n_features, actor_actions, lr, beta, cuda=False):
        self.name = name
        self.num_outs = actor_actions
        self.net = NeuralNet(n_features, actor_actions, b_actor=True)
        self.optimizer = Adam(self.net.parameters(), lr=lr)
        self.cuda = cuda
        self.beta = beta
        self.losses = []
        # device sampler stream: keyed from torch's seed WITHOUT drawing from the generator (a draw here
        # would shift the initial weights of every network constructed afterwards vs the reference)
        ActorNetwork._instances += 1
        self._seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + ActorNetwork._instances) & (2 ** 64 - 1)
        self._calls = 0
        self._replay = None
        self.last_probs = None

    def replay(self, actions):
        """Inject recorded action samples (an iterable of per-call tensors/arrays): subsequent
        ``sample_action`` calls return them instead of the device sampler's draw (parity under replay)."""
        self._replay = iter(actions) if actions is not None else None

    def _sample_small_host(self, obs):
        """Few rows of host observations (the a2c_org_test.py loop: one env, one call per step): the kernel reads the rows from
        and writes the actions to pinned host memory, so the call is one launch + one stream sync — no copies, no allocations."""
        lib = _lib.load()
        x = np.asarray(obs, dtype=np.float32)
        F_ = self.net.state_dim
        rows = x.size // F_
        st = self.__dict__.get("_small")
        if st is None or st["rows"] < rows:
            cap = max(rows, 16)
            hx = torch.zeros(cap * F_, dtype=torch.float32).pin_memory()
            ha = torch.zeros(cap, dtype=torch.int64).pin_memory()
            st = self._small = dict(rows=cap, hx=hx, ha=ha, nx=hx.numpy(), na=ha.numpy(),
                                    probs=torch.empty(cap, self.num_outs, dtype=torch.float32, device=self.net.flat.device))
        st["nx"][:rows * F_] = x.reshape(-1)
        with torch.cuda.device(self.net.flat.device):
            stream = torch.cuda.current_stream()
            _lib.check(lib.ia2c_actor_sample(_lib.ptr(self.net.flat.detach()), st["hx"].data_ptr(), None, st["ha"].data_ptr(),
                                             _lib.ptr(st["probs"]), rows, F_, self.num_outs, self._seed, self._calls,
                                             stream.cuda_stream), "ia2c_actor_sample")
            stream.synchronize()
        self._calls += 1
        self.last_probs = st["probs"][:rows]
        return torch.from_numpy(st["na"][:rows].copy())

    def sample_action(self, obs, grad=False):
        lib = _lib.load()
        dev = self.net.flat.device
        in_dev = obs.device if isinstance(obs, torch.Tensor) else torch.device("cpu")
        if in_dev.type == "cpu" and not _is_index_input(obs):
            x_h = obs.detach().numpy() if isinstance(obs, torch.Tensor) else np.asarray(obs)
            if x_h.ndim >= 1 and x_h.shape[-1] == self.net.state_dim and x_h.size <= 64 * self.net.state_dim:
                lead = x_h.shape[:-1]
                actions = self._sample_small_host(x_h)
                if self._replay is not None:
                    return torch.as_tensor(np.asarray(next(self._replay))).reshape(lead).to(torch.int64)
                return actions.reshape(lead)
        if _is_index_input(obs):   # class indices standing for one-hot rows (index fast path)
            idx, x2 = _checked_indices(obs, self.net.state_dim, dev)
            lead, entry = idx.shape, lib.ia2c_actor_sample_index
        else:
            x = torch.as_tensor(obs).to(dev, torch.float32)
            lead, entry = x.shape[:-1], lib.ia2c_actor_sample
            x2 = x.reshape(-1, self.net.state_dim).contiguous()
        rows = x2.shape[0]
        actions = torch.empty(rows, dtype=torch.int64, device=dev)
        probs = torch.empty(rows, self.num_outs, dtype=torch.float32, device=dev)
        _lib.check(entry(_lib.ptr(self.net.flat.detach()), _lib.ptr(x2), None, _lib.ptr(actions),
                         _lib.ptr(probs), rows, self.net.state_dim, self.num_outs, self._seed,
                         self._calls, _lib.stream_ptr()), "ia2c_actor_sample")
        self._calls += 1
        self.last_probs = probs
        if self._replay is not None:
            return torch.as_tensor(np.asarray(next(self._replay))).reshape(lead).to(in_dev, torch.int64)
        return actions.reshape(lead).to(in_dev)

    def action_distribution(self, obs, grad=False):
        in_dev = obs.device if isinstance(obs, torch.Tensor) else torch.device("cpu")
        if not grad:
            with torch.no_grad():
                out = self.net(obs)
        else:
            out = self.net(obs)
        return out.to(in_dev)

    def batch_update(self, obs, act, adv, retain=False):
        dev = self.net.flat.device
        if _no_graph(adv) and _no_graph(obs):   # gradient-free advantage (ia2c.py:127): one fused C call
            adv_f = adv
            if isinstance(adv, torch.Tensor) and adv.dim() >= 2 and adv.shape[-1] == 1:
                adv_f = adv.squeeze(-1)
            fused = _fused_update(self.net, self.optimizer, 1, obs, act, adv_f, self.beta, "ActorNetwork.batch_update")
            if fused is not None:
                _raise_on_status(fused[1], self.net.state_dim)
                self.losses.append(fused[0])
                if len(self.losses) > 20:
                    del self.losses[0]
                self.actor_loss = np.mean(self.losses)
                return
        probs = self.net.forward(obs)
        idx = _index_like(torch.as_tensor(act), probs.shape[:-1], "ActorNetwork.batch_update").to(dev)
        adv_d = adv.to(dev)  # differentiable copy (the advantage may carry gradient, Q7)
        if adv_d.dim() == probs.dim() and adv_d.shape[-1] == 1:
            adv_d = adv_d.squeeze(-1)
        loss = _ActorLossFunction.apply(probs, idx, adv_d, self.beta)
        loss.backward(retain_graph=retain)      # NO zero_grad: gradients accumulate across updates (Q2)
        self.optimizer.step()
        get_loss = loss.detach().cpu().numpy()
        if int(_ActorLossFunction.last_status.item()):  # the reference's Categorical validation raises here
            raise ValueError("Expected parameter probs of distribution Categorical to satisfy the constraint Simplex()")
        self.losses.append(get_loss)
        if len(self.losses) > 20:
            del self.losses[0]
        self.actor_loss = np.mean(self.losses)
