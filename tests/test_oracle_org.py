"""Oracle pin: oracle/org.py against the truth table and walk recorded from the real Org class."""
import numpy as np

from oracle import org as O


def test_truth_table_closed_form_and_rules(golden):
    g = golden("org_table.npz")
    tab = g["table"]
    s, a, r, prev = tab[:, 0].astype(int), tab[:, 1].astype(int), tab[:, 2], tab[:, 3].astype(int)
    s2, r2, obs = tab[:, 4].astype(int), tab[:, 5], tab[:, 6:12]
    ns, nr, cls = O.org_step_joint(s, r, a)
    assert np.array_equal(ns, s2)
    assert np.array_equal(nr, r2)  # bit-exact fp64
    assert np.array_equal(O.make_obs(prev, cls, np.float64), obs)
    assert not tab[:, 12].any()  # done is always False (Q5)
    for i in range(len(tab)):
        st, rw = O.org_step_scalar(int(s[i]), float(r[i]), int(a[i]))
        assert st == s2[i] and rw == r2[i]


def test_random_walk(golden):
    g = golden("org_table.npz")
    env = O.OrgBatchRef(1)
    assert np.array_equal(env.reset()[0], g["reset_obs"].astype(np.float32))
    for t, a in enumerate(g["walk_actions"]):
        obs, r, _ = env.step_joint(np.array([a]))
        assert env.state[0] == g["walk_state"][t]
        assert r[0] == g["walk_reward"][t]
        assert np.array_equal(obs[0].astype(np.float64), g["walk_obs"][t])


def test_org_n_reduces_to_reference_at_two_agents():
    rng = np.random.RandomState(0)
    s = rng.randint(0, 5, 5000)
    r = rng.randn(5000) * 50
    acts = rng.randint(0, 3, (5000, 2))
    a = O.org_step_agents(s, r, acts)
    b = O.org_step_joint(s, r, acts[:, 0] * 3 + acts[:, 1])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_time_limit_autoreset():
    env = O.OrgBatchRef(3, max_episode_steps=4)
    env.reset()
    for t in range(4):
        obs, r, trunc = env.step_joint(np.array([8, 0, 4]))
    assert trunc.all()
    assert np.array_equal(obs, np.tile(O.RESET_OBS.astype(np.float32), (3, 1)))  # reset obs returned (Q14)
    assert r[1] == -100 + (-100 + (-100 + 6 / 10) / 10) / 10  # reward is the real pre-reset reward
    assert (env.reward == 0).all() and (env.state == 2).all()
