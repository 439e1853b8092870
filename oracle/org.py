"""CPU restatement of the Org domain.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows the reference's ``Org`` class:
  * state transition + reward recurrence ........ Org.py:51-104
  * observation class of a state ................ Org.py:13-20
  * observation memory shift (MEM, MEM_SIZE=1) .. Org.py:105-114
  * reset ....................................... Org.py:128-148
  * unknown action codes leave state AND reward untouched but still shift the
    observation memory (no ``else`` branch at Org.py:52-104)

Two forms are given and tested against each other and against the golden truth
table recorded from the real class (tests/golden/org_table.npz):
``org_step_scalar`` restates the branch structure as a rule table, and
``org_step_joint`` / ``org_step_agents`` are the vectorised closed form that the
CUDA kernel implements.

Pinned against: tests/golden/org_table.npz (truth table + 4000-step walk of the
real class) and the env tapes inside tests/golden/ia2c_*.npz, a2c_org.npz.
"""
from __future__ import annotations

import numpy as np

N_STATES = 5
RESET_STATE = 2
RESET_OBS = np.array([0.0, 1.0, 0.0, 0.0, 1.0, 0.0])

# joint action -> (movement rule, base reward when the new state is not 0)        Org.py:52-104
_RULES = {
    0: ("down", 6.0), 1: ("down", 1.0), 3: ("down", 1.0),
    2: ("stay", 1.0), 6: ("stay", 1.0), 4: ("stay", 5.0),
    5: ("up1", 1.0), 7: ("up1", 1.0),
    8: ("up2", 1.0),
}


def obs_class(state):
    """Org.py:13-20 — 0 for states {0,1}, 1 for {2,3}, 2 for {4}."""
    state = np.asarray(state)
    return np.where(state < 2, 0, np.where(state < 4, 1, 2))


def org_step_scalar(state: int, reward: float, action: int):
    """One env, one step, rule-table form.  Returns (state', reward')."""
    rule = _RULES.get(int(action))
    if rule is None:
        return state, reward
    move, base = rule
    if move == "down":
        if state <= 1:
            return 0, -100 + reward / 10
        return state - 1, base + reward / 10
    if move == "stay":
        if state == 0:
            return state, -100 + reward / 10
        return state, base + reward / 10
    if move == "up1":
        return min(state + 1, 4), base + reward / 10
    # up2
    return min(state + 2, 4), base + reward / 10


def _finish(state, reward, prev_cls, new_state, base, valid):
    new_state = np.where(valid, new_state, state)
    stepped = base.astype(np.float64) + reward / 10.0  # IEEE divide then add (Q17)
    new_reward = np.where(valid, stepped, reward)
    cls = obs_class(new_state)
    return new_state.astype(np.int32), new_reward, cls.astype(np.int32)


def org_step_joint(state, reward, joint):
    """Vectorised closed form for the reference's 2-agent joint code 0..8 (others: no-op).

    state int[E], reward f64[E], joint int[E] -> (state', reward', cls') ; Appendix A.1.
    """
    state = np.asarray(state, dtype=np.int64)
    reward = np.asarray(reward, dtype=np.float64)
    joint = np.asarray(joint, dtype=np.int64)
    valid = (joint >= 0) & (joint <= 8)
    j = np.where(valid, joint, 4)
    a1, a2 = j // 3, j % 3
    n_s = (a1 == 0).astype(np.int64) + (a2 == 0)
    n_g = (a1 == 2).astype(np.int64) + (a2 == 2)
    delta = np.maximum(-1, n_g - n_s)
    new_state = np.clip(state + delta, 0, 4)
    base = np.where(new_state == 0, -100.0, np.where(j == 0, 6.0, np.where(j == 4, 5.0, 1.0)))
    return _finish(state, reward, None, new_state, base, valid)


def org_step_agents(state, reward, actions):
    """Org-N (builder-defined generalisation): actions int[E,N] in {0,1,2}.

    Reduces exactly to ``org_step_joint`` at N=2 (tested).
    """
    state = np.asarray(state, dtype=np.int64)
    reward = np.asarray(reward, dtype=np.float64)
    actions = np.asarray(actions, dtype=np.int64)
    n = actions.shape[1]
    n_s = (actions == 0).sum(1)
    n_b = (actions == 1).sum(1)
    n_g = (actions == 2).sum(1)
    delta = np.clip(n_g - n_s, -1, 2)
    new_state = np.clip(state + delta, 0, 4)
    base = np.where(new_state == 0, -100.0, np.where(n_s == n, 6.0, np.where(n_b == n, 5.0, 1.0)))
    valid = np.ones_like(state, dtype=bool)
    return _finish(state, reward, None, new_state, base, valid)


def make_obs(prev_cls, cls, dtype=np.float32):
    """[onehot3(prev class), onehot3(new class)]  (Org.py:112-114)."""
    prev_cls = np.asarray(prev_cls)
    cls = np.asarray(cls)
    eye = np.eye(3, dtype=dtype)
    return np.concatenate([eye[prev_cls], eye[cls]], axis=-1)


class OrgBatchRef:
    """E independent Org envs with optional TimeLimit + same-step autoreset (Q14)."""

    def __init__(self, num_envs, max_episode_steps=None):
        self.E = num_envs
        self.max_episode_steps = max_episode_steps
        self.reset()

    def reset(self):
        self.state = np.full(self.E, RESET_STATE, dtype=np.int32)
        self.reward = np.zeros(self.E, dtype=np.float64)
        self.cls = obs_class(self.state).astype(np.int32)  # newest obs class
        self.prev_cls = np.ones(self.E, dtype=np.int32)     # memory slot literal [0,1,0] (Org.py:37,138)
        self.elapsed = np.zeros(self.E, dtype=np.int64)
        return make_obs(self.prev_cls, self.cls)

    def _after(self, new_state, new_reward, new_cls):
        self.prev_cls, self.cls = self.cls, new_cls
        self.state, self.reward = new_state, new_reward
        out_reward = new_reward.copy()
        self.state_pre_reset = new_state.copy()
        self.elapsed += 1
        trunc = np.zeros(self.E, dtype=bool)
        if self.max_episode_steps is not None:
            trunc = self.elapsed >= self.max_episode_steps
            if trunc.any():
                self.state = np.where(trunc, RESET_STATE, self.state).astype(np.int32)
                self.reward = np.where(trunc, 0.0, self.reward)
                self.cls = np.where(trunc, 1, self.cls).astype(np.int32)
                self.prev_cls = np.where(trunc, 1, self.prev_cls).astype(np.int32)
                self.elapsed = np.where(trunc, 0, self.elapsed)
        return make_obs(self.prev_cls, self.cls), out_reward, trunc

    def step_joint(self, joint):
        return self._after(*org_step_joint(self.state, self.reward, joint))

    def step_agents(self, actions):
        return self._after(*org_step_agents(self.state, self.reward, actions))
