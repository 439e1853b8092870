"""GPU parity: Org env kernels (through the C ABI) vs the oracle and the golden truth table.  Bit-exact."""
import numpy as np
import pytest

from oracle import org as O
from tests.helpers import dev, dptr, host

pytestmark = pytest.mark.gpu


def _run_joint(s, r, prev, cur, joint, max_steps=0, elapsed=None):
    import torch
    from ia2c_b200 import _lib

    lib = _lib.load()
    E = len(s)
    state, hist = dev(s, torch.int32), dev(r, torch.float64)
    cls = dev(np.stack([prev, cur], 1), torch.uint8)
    el = dev(np.zeros(E) if elapsed is None else elapsed, torch.int32)
    obs = torch.empty(E, 6, device="cuda")
    rew = torch.empty(E, dtype=torch.float64, device="cuda")
    rew32 = torch.empty(E, device="cuda")
    trace = torch.empty(E, dtype=torch.int32, device="cuda")
    trunc = torch.empty(E, dtype=torch.uint8, device="cuda")
    _lib.check(lib.ia2c_org_step_joint(_lib.ptr(state), _lib.ptr(hist), _lib.ptr(cls), _lib.ptr(el), dptr(joint, torch.int32),
                                       _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(rew32), _lib.ptr(trace), _lib.ptr(trunc), E, max_steps,
                                       _lib.stream_ptr()))
    torch.cuda.synchronize()
    return host(state), host(hist), host(obs), host(rew), host(rew32), host(trace), host(trunc), host(cls)


def test_truth_table_bit_exact(golden):
    tab = golden("org_table.npz")["table"]
    s, a, r, prev = tab[:, 0].astype(int), tab[:, 1].astype(int), tab[:, 2], tab[:, 3].astype(int)
    cur = O.obs_class(s)  # irrelevant to the result except through the memory shift: table rows set obs[3:6]=prev
    state, hist, obs, rew, rew32, trace, trunc, cls = _run_joint(s, r, cur, prev, a)
    assert np.array_equal(state, tab[:, 4].astype(np.int32))
    assert np.array_equal(rew, tab[:, 5]) and np.array_equal(hist, tab[:, 5])
    assert np.array_equal(obs.astype(np.float64), tab[:, 6:12])
    assert np.array_equal(rew32, tab[:, 5].astype(np.float32))
    assert not trunc.any()


def test_walk_vec_env_and_drop_in_class(golden):
    import torch
    from ia2c_b200.org_env import Org, OrgVecEnv

    g = golden("org_table.npz")
    acts = g["walk_actions"][:600]
    env = OrgVecEnv(1)
    single = Org()
    o0, _ = single.reset()
    assert np.array_equal(o0, g["reset_obs"])
    for t, a in enumerate(acts):
        obs, r, term, trunc, _ = env.step(np.array([a]))
        assert obs.dtype == np.float32 and r.dtype == np.float64
        assert np.array_equal(obs[0].astype(np.float64), g["walk_obs"][t]) and r[0] == g["walk_reward"][t]
        o, rr, d1, d2, info = single.step(int(a))
        assert o is single.observation                      # aliasing quirk Q4
        assert np.array_equal(o, g["walk_obs"][t]) and rr == g["walk_reward"][t] and single.state == g["walk_state"][t]
        assert d1 is False and d2 is False and info == {}
    # numpy 0-d / torch 0-d actions as the scripts pass them
    o, rr, *_ = single.step(np.array(4))
    o, rr, *_ = single.step(torch.tensor(8))


@pytest.mark.parametrize("E,N", [(e, n) for e in (1, 31, 33, 1000, 4096) for n in (2, 3, 8, 9, 33, 64, 256)] +
                         # the 32-envs-per-warp kernel (N % 4 == 0): several envs per load (N <= 64), masked tail words, 1 / 2 / 4 / 8 words per lane
                         [(e, n) for e in (1, 33, 1000) for n in (12, 36, 100, 128, 260, 516, 1020)])
def test_org_n_matches_oracle(E, N):
    import torch
    from ia2c_b200.org_env import OrgVecEnv

    rng = np.random.RandomState(E * 1000 + N)
    env = OrgVecEnv(E, n_agents=N, max_episode_steps=7)
    ref = O.OrgBatchRef(E, max_episode_steps=7)
    obs0, _ = env.reset()
    assert obs0.dtype == np.float32 and np.array_equal(obs0, ref.reset())
    p = rng.dirichlet([1, 1, 1])
    for t in range(20):
        a = rng.choice(3, size=(E, N), p=p).astype(np.uint8)
        if t % 5 == 4:
            a[:] = rng.randint(0, 3)  # unanimous steps exercise the 6 / 5 base rewards
        obs, r, term, trunc, _ = env.step(torch.as_tensor(a).cuda())
        robs, rr, rtrunc = ref.step_agents(a)
        assert np.array_equal(host(obs), robs)
        assert np.array_equal(host(r), rr)
        assert np.array_equal(host(trunc), rtrunc)
        assert np.array_equal(host(env.state), ref.state) and np.array_equal(host(env.hist), ref.reward)
        assert np.array_equal(host(env.state_trace), ref.state_pre_reset)


def test_joint_equals_agents_at_two_and_large_batch_properties():
    import torch
    from ia2c_b200.org_env import OrgVecEnv

    E = 1 << 20
    rng = np.random.RandomState(3)
    a = rng.randint(0, 3, size=(E, 2)).astype(np.uint8)
    e1, e2 = OrgVecEnv(E), OrgVecEnv(E)
    ref = O.OrgBatchRef(E)
    for t in range(6):
        a = rng.randint(0, 3, size=(E, 2)).astype(np.uint8)
        o1, r1, *_ = e1.step(torch.as_tensor(a).cuda())
        o2, r2, *_ = e2.step(torch.as_tensor(a[:, 0].astype(np.int32) * 3 + a[:, 1]).cuda())
        assert torch.equal(o1, o2) and torch.equal(r1, r2)
        robs, rr, _ = ref.step_agents(a)
        assert np.array_equal(host(o1), robs) and np.array_equal(host(r1), rr)
    # size-independent invariants: one-hot pairs, state range, reward bounds of r <- base + r/10
    assert torch.all(o1.sum(1) == 2) and torch.all((e1.state >= 0) & (e1.state <= 4))
    assert float(e1.hist.max()) < 6.7 and float(e1.hist.min()) > -111.2


def test_bad_arguments_fail_loudly():
    from ia2c_b200 import _lib
    lib = _lib.load()
    assert lib.ia2c_org_reset(None, None, None, None, None, 4, None) == -1
    assert b"ia2c_org_reset" in lib.ia2c_last_error()
