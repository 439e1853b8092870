"""Minimal stand-in for the parts of gymnasium 0.29.1 that the reference touches.

TEST INFRASTRUCTURE (see oracle/__init__.py).  gymnasium is pinned by the
reference (requirements.txt:37) but is not installed in the build image and
cannot be installed (no network).  The reference uses exactly:

  * ``gym.Env`` as a base class                       (Org.py:12)
  * ``gym.spaces.Discrete`` / ``gym.spaces.Box``       (Org.py:27-28,40)
  * ``gymnasium.envs.registration.register``           (ia2c.py:33-38)
  * ``gym.make_vec(id, num_envs=E)``                   (ia2c.py:42)

``make_vec`` here is a synchronous in-process vector env that reproduces the
three behaviours of gymnasium 0.29.1's vector envs that reach the numbers:
TimeLimit truncation after ``max_episode_steps``, *same-step* autoreset (the
observation returned on the truncating step is the reset observation while the
reward is the real pre-reset reward), and dtype handling
(observations concatenated into the Box dtype float32, rewards float64).
It contributes no arithmetic.

``install()`` registers this module as ``gymnasium`` in ``sys.modules`` so that
the unmodified reference files import it.
"""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np


class Env:
    metadata: dict = {}

    def reset(self, seed=None, options=None):  # pragma: no cover - interface
        raise NotImplementedError

    def step(self, action):  # pragma: no cover - interface
        raise NotImplementedError


class Discrete:
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = int(start)
        self.shape = ()
        self.dtype = np.dtype(np.int64)


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        low = np.asarray(low)
        high = np.asarray(high)
        if shape is None:
            shape = low.shape
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(low, self.shape).astype(self.dtype)
        self.high = np.broadcast_to(high, self.shape).astype(self.dtype)


_REGISTRY: dict[str, dict] = {}


def register(id, entry_point=None, max_episode_steps=None, **kwargs):
    _REGISTRY[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps, kwargs=kwargs)


def _load(entry_point):
    if callable(entry_point):
        return entry_point
    mod, attr = entry_point.split(":")
    return getattr(importlib.import_module(mod), attr)


class SyncVectorEnv:
    """E independent copies stepped in a Python loop, TimeLimit + same-step autoreset."""

    def __init__(self, env_fns, max_episode_steps=None):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.max_episode_steps = max_episode_steps
        self._elapsed = np.zeros(self.num_envs, dtype=np.int64)
        space = getattr(self.envs[0], "observation_space", None)
        self._obs_dtype = getattr(space, "dtype", np.dtype(np.float32))
        self.single_observation_space = space
        self.single_action_space = getattr(self.envs[0], "action_space", None)

    def reset(self, seed=None, options=None):
        obs = []
        for env in self.envs:
            o, _ = env.reset()
            obs.append(np.array(o, dtype=self._obs_dtype))
        self._elapsed[:] = 0
        return np.stack(obs), {}

    def step(self, actions):
        obs, rew, term, trunc = [], [], [], []
        # env-internal state right after the step, before any autoreset (for golden tapes)
        self.last_pre_reset_state = np.zeros(self.num_envs, dtype=np.int64)
        for i, (env, a) in enumerate(zip(self.envs, actions)):
            o, r, te, tr, _ = env.step(a)
            self.last_pre_reset_state[i] = getattr(env, "state", -1)
            self._elapsed[i] += 1
            if self.max_episode_steps is not None and self._elapsed[i] >= self.max_episode_steps:
                tr = True
            if te or tr:
                o, _ = env.reset()
                self._elapsed[i] = 0
            obs.append(np.array(o, dtype=self._obs_dtype))
            rew.append(r)
            term.append(bool(te))
            trunc.append(bool(tr))
        return (np.stack(obs), np.array(rew, dtype=np.float64), np.array(term, dtype=np.bool_),
                np.array(trunc, dtype=np.bool_), {})

    def close(self):
        pass


def make_vec(id, num_envs=1, **kwargs):
    spec = _REGISTRY[id]
    cls = _load(spec["entry_point"])
    return SyncVectorEnv([cls for _ in range(num_envs)], max_episode_steps=spec["max_episode_steps"])


def make(id, **kwargs):  # unused by the reference; here for completeness
    spec = _REGISTRY[id]
    return _load(spec["entry_point"])()


def install():
    """Expose this module as ``gymnasium`` (+ ``.spaces``, ``.envs.registration``)."""
    if "gymnasium" in sys.modules and getattr(sys.modules["gymnasium"], "__standin__", False):
        return sys.modules["gymnasium"]
    gym = types.ModuleType("gymnasium")
    gym.__standin__ = True
    gym.Env = Env
    gym.make_vec = make_vec
    gym.make = make
    gym.register = register
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Discrete = Discrete
    spaces.Box = Box
    gym.spaces = spaces
    envs = types.ModuleType("gymnasium.envs")
    registration = types.ModuleType("gymnasium.envs.registration")
    registration.register = register
    envs.registration = registration
    gym.envs = envs
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = spaces
    sys.modules["gymnasium.envs"] = envs
    sys.modules["gymnasium.envs.registration"] = registration
    return gym
