"""Org domain on the GPU: the batched vector env and the drop-in single-env ``Org`` class.

Reference interface mirrored here (thinclab/IA2C):
  * ``Org`` (gym.Env): attributes state/done/reward/hist/action_space/observation_space/observation,
    ``reset(seed=None, options={}) -> (observation, {})``, ``step(action) -> (observation, reward,
    done, done, {})``, ``render(mode)``                                   Org.py:12-151
  * ``gym.make_vec("Org-v0", num_envs=E)`` semantics (TimeLimit 30, same-step autoreset, float32
    observations [E,6], float64 rewards [E])                               ia2c.py:33-42,72,85

All arithmetic runs in ``ia2c_org_*`` kernels (csrc/org_env.cu); there is no CPU implementation here.
"""
from __future__ import annotations

import numpy as np

from . import _lib

try:  # the reference subclasses gymnasium.Env; use it when installed, otherwise a structural stand-in
    import gymnasium as _gym

    _EnvBase = _gym.Env
    _Discrete = _gym.spaces.Discrete
    _Box = _gym.spaces.Box
except Exception:  # gymnasium is not part of this image
    class _EnvBase:  # noqa: D401
        """Structural stand-in for gymnasium.Env (no behaviour)."""

    class _Discrete:
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)

    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape if shape is None else tuple(shape)
            self.dtype = np.dtype(dtype)


N_FEATURES = 6
RESET_STATE = 2


class OrgVecEnv:
    """E independent Org instances (x N agents) stepped by one kernel launch.

    ``step`` accepts either the reference's joint action codes ``int[E]`` (0..8, two agents; other codes are
    no-ops that still shift the observation memory) or per-agent actions ``int[E,N]`` in {0,1,2} (Org-N).
    CUDA tensors in -> CUDA tensors out (no host sync); numpy / CPU tensors in -> numpy out.
    """

    def __init__(self, num_envs, n_agents=2, max_episode_steps=None, device=None):
        torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.num_envs = int(num_envs)
        self.n_agents = int(n_agents)
        self.max_episode_steps = int(max_episode_steps) if max_episode_steps else 0
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        E = self.num_envs
        dev = self.device
        self.state = torch.empty(E, dtype=torch.int32, device=dev)
        self.hist = torch.empty(E, dtype=torch.float64, device=dev)
        self.cls = torch.empty(E, 2, dtype=torch.uint8, device=dev)
        self.elapsed = torch.zeros(E, dtype=torch.int32, device=dev)
        self.obs = torch.empty(E, N_FEATURES, dtype=torch.float32, device=dev)
        self.reward = torch.zeros(E, dtype=torch.float64, device=dev)
        self.reward_f32 = torch.zeros(E, dtype=torch.float32, device=dev)
        self.state_trace = torch.empty(E, dtype=torch.int32, device=dev)
        self.truncated = torch.zeros(E, dtype=torch.uint8, device=dev)
        self.single_observation_space = _Box(-np.ones(6), np.ones(6))
        self.single_action_space = _Discrete(2)
        self.reset_device()

    def reset_device(self):
        """Reset all envs; returns the device observation tensor f32[E,6] (no host sync)."""
        with self._guard():
            _lib.check(self.lib.ia2c_org_reset(_lib.ptr(self.state), _lib.ptr(self.hist), _lib.ptr(self.cls),
                                               _lib.ptr(self.elapsed), _lib.ptr(self.obs), self.num_envs,
                                               _lib.stream_ptr()), "ia2c_org_reset")
        return self.obs

    def reset(self, seed=None, options=None):
        """gymnasium vector-env signature: (float32 numpy observations [E,6], info)."""
        return self.reset_device().cpu().numpy(), {}

    def _guard(self):
        import torch

        return torch.cuda.device(self.device)

    def step_device(self, actions):
        """actions: CUDA int32[E] joint codes or uint8[E,N] per-agent.  Returns device tensors (views of
        internal buffers, overwritten by the next step): obs f32[E,6], reward f64[E], truncated u8[E]."""
        with self._guard():
            args = (_lib.ptr(self.state), _lib.ptr(self.hist), _lib.ptr(self.cls), _lib.ptr(self.elapsed))
            outs = (_lib.ptr(self.obs), _lib.ptr(self.reward), _lib.ptr(self.reward_f32), _lib.ptr(self.state_trace),
                    _lib.ptr(self.truncated))
            if actions.dim() == 1:
                _lib.check(self.lib.ia2c_org_step_joint(*args, _lib.ptr(actions), *outs, self.num_envs,
                                                        self.max_episode_steps, _lib.stream_ptr()), "ia2c_org_step_joint")
            else:
                _lib.check(self.lib.ia2c_org_step_agents(*args, _lib.ptr(actions), *outs, self.num_envs,
                                                         actions.shape[1], self.max_episode_steps, _lib.stream_ptr()),
                           "ia2c_org_step_agents")
        return self.obs, self.reward, self.truncated

    def step(self, actions):
        import torch

        on_device = isinstance(actions, torch.Tensor) and actions.is_cuda
        a = torch.as_tensor(np.asarray(actions) if not isinstance(actions, torch.Tensor) else actions)
        if a.dim() == 0:
            a = a.reshape(1)
        if a.shape[0] != self.num_envs:
            raise ValueError(f"expected actions for {self.num_envs} envs, got shape {tuple(a.shape)}")
        if a.dim() == 1:
            a = a.to(self.device, torch.int32).contiguous()
        elif a.dim() == 2:
            if a.shape[1] != self.n_agents:
                raise ValueError(f"expected {self.n_agents} agent actions per env, got {a.shape[1]}")
            a = a.to(self.device, torch.uint8).contiguous()
        else:
            raise ValueError("actions must be [E] joint codes or [E,N] per-agent actions")
        obs, rew, trunc = self.step_device(a)
        if on_device:
            term = torch.zeros_like(trunc, dtype=torch.bool)
            return obs, rew, term, trunc.bool(), {}
        obs_h, rew_h, trunc_h = obs.cpu().numpy(), rew.cpu().numpy(), trunc.cpu().numpy().astype(np.bool_)
        return obs_h, rew_h, np.zeros(self.num_envs, dtype=np.bool_), trunc_h, {}

    def close(self):
        pass


class Org(_EnvBase):
    """Drop-in for the reference's ``Org`` (default flags MEM=True, MEM_SIZE=1, STATE_VISIBLE=False).

    One env instance backed by E=1 device arrays.  ``observation`` is a float64 numpy array that ``step``
    updates IN PLACE and returns by reference, exactly like the reference (a2c_org_test.py relies on this
    aliasing: ``states is next_states``).
    """

    def __init__(self):
        self._vec = OrgVecEnv(1, n_agents=2, max_episode_steps=None)
        torch = _lib.require_cuda()
        self._torch = torch
        # one pinned (host-mapped) line the step kernel reads the action from and writes its outputs to — a single env is
        # pure launch latency, so there is no H2D / D2H copy and no torch op per step: action i32 @0, state i32 @4,
        # reward f64 @8, observation f32[6] @16
        self._h = torch.zeros(64, dtype=torch.uint8).pin_memory()
        self._np = self._h.numpy()
        self._np_action, self._np_state = self._np[0:4].view(np.int32), self._np[4:8].view(np.int32)
        self._np_reward, self._np_obs = self._np[8:16].view(np.float64), self._np[16:40].view(np.float32)
        self._h_ptr = self._h.data_ptr()
        self.done = False
        self.hist = 0
        self.action_space = _Discrete(2)          # Org.py:27 (sic)
        low = np.array([-1.0] * 6)
        high = np.array([1.0] * 6)
        self.observation_space = _Box(low, high)  # Org.py:38-40
        self.observation = np.array([0.0, 1.0, 0.0, 0.0, 1.0, 0.0])
        self._synced_obs = self.observation.copy()
        self._state = RESET_STATE
        self._reward = 0

    # state / reward mirror the device arrays; assigning them pushes to the device (tests poke them)
    @property
    def state(self):
        return self._state

    @state.setter
    def state(self, v):
        self._state = int(v)
        self._vec.state.fill_(int(v))

    @property
    def reward(self):
        return self._reward

    @reward.setter
    def reward(self, v):
        self._reward = v
        self._vec.hist.fill_(float(v))

    def _push_observation(self):
        if np.array_equal(self.observation, self._synced_obs):
            return
        prev = int(np.argmax(self.observation[0:3]))
        cur = int(np.argmax(self.observation[3:6]))
        self._vec.cls.copy_(self._torch.tensor([[prev, cur]], dtype=self._torch.uint8))
        self._synced_obs[:] = self.observation

    def step(self, action):
        try:
            code = int(action)
        except Exception:
            code = int(np.asarray(action).reshape(-1)[0])
        self._push_observation()  # honour in-place edits of .observation by the caller
        v, h = self._vec, self._h_ptr
        self._np_action[0] = code
        with v._guard():
            stream = self._torch.cuda.current_stream()
            _lib.check(v.lib.ia2c_org_step_joint(_lib.ptr(v.state), _lib.ptr(v.hist), _lib.ptr(v.cls), None, h, h + 16, h + 8, None,
                                                 h + 4, None, 1, 0, stream.cuda_stream), "ia2c_org_step_joint")
            stream.synchronize()
        self.observation[:] = self._np_obs        # in place: the returned array is the same object (Q4)
        self._synced_obs[:] = self._np_obs
        valid = 0 <= code <= 8
        self._reward = float(self._np_reward[0]) if valid else self._reward   # untouched on unknown codes (Q16)
        self._state = int(self._np_state[0])
        return (self.observation, self._reward, self.done, self.done, {})

    def reset(self, seed=None, options={}):
        self._vec.reset_device()
        self._state = RESET_STATE
        self.done = False
        self._reward = 0
        self.observation = np.array([0.0, 1.0, 0.0, 0.0, 1.0, 0.0])  # a NEW array, as in the reference
        self._synced_obs = self.observation.copy()
        return (self.observation, {})

    def getObsFromState(self):
        return 0 if self._state < 2 else (1 if self._state < 4 else 2)

    def render(self, mode=None):
        print(self._state)
