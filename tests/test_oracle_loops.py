"""Oracle pin: oracle/loops.py replayed against unmodified-reference runs of ia2c.py and a2c_org_test.py."""
import numpy as np
import pytest

from oracle import loops as L

RTOL = 1e-5


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def ia2c_state_from_golden(g, dtype=np.float32):
    lr_c, lr_a, beta, gamma = g["meta_hyper"]
    return L.IA2CState(
        actor=np.stack([g["act1/init"][0], g["act2/init"][0]]).astype(dtype),
        critic=np.stack([g["crit1/init"][0], g["crit2/init"][0]]).astype(dtype),
        filter_action=np.stack([g["bf0/filterAction"][0], g["bf1/filterAction"][0]]),
        lr_c=lr_c, lr_a=lr_a, beta=beta, gamma=gamma, T=int(g["meta_T"]))


def ia2c_tapes(g, ep):
    T = int(g["meta_T"])
    sl = slice(ep * (T + 1), (ep + 1) * (T + 1))
    actions = np.stack([g["act1/sampled"][sl], g["act2/sampled"][sl]], axis=-1)      # [T+1,E,2]
    u = np.stack([g["bf0/u"][sl, :, 0], g["bf1/u"][sl, :, 0]], axis=-1)[..., None]    # [T+1,E,2,K=1]
    return sl, actions, u


@pytest.mark.parametrize("name", ["ia2c_E10.npz", "ia2c_E64.npz"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ia2c_replay(golden, name, dtype):
    g = golden(name)
    T, E = int(g["meta_T"]), int(g["meta_n_envs"])
    st = ia2c_state_from_golden(g, dtype)
    for ep in range(int(g["meta_episodes"])):
        sl, actions, u = ia2c_tapes(g, ep)
        traj, upd = L.ia2c_episode(st, E, actions=actions, u_belief=u)
        es = slice(ep * T, (ep + 1) * T)
        # bit-exact: env state, reward (fp64), observations
        assert np.array_equal(traj["state"], g["env/state_pre_reset"][es])
        assert np.array_equal(traj["reward"], g["env/reward"][es])
        assert np.array_equal(traj["obs"][1:], g["env/obs"][es])
        assert np.array_equal(traj["obs"][0], g["env/reset_obs"][ep])
        assert np.array_equal(traj["ep_return"], g["reward_lst"][ep])
        # bit-exact: beliefs and predicted actions (bf0 = agent 1's filter over agent 2)
        for i, f in enumerate(("bf0", "bf1")):
            assert np.array_equal(traj["belief"][:, :, i, 0], g[f"{f}/bprime"][sl])
            assert np.array_equal(traj["pred"][:, :, i, 0], g[f"{f}/ap"][sl])
        # update inputs as the reference's batch_update saw them
        assert np.array_equal(traj["obs"][:T], g["crit1/upd_obs"][ep])
        for i, (c, a) in enumerate((("crit1", "act1"), ("crit2", "act2"))):
            assert rel_err(upd["critic_target"][i], g[f"{c}/upd_target"][ep][..., 0]) < RTOL
            assert rel_err(upd["critic_loss"][i], g[f"{c}/upd_loss"][ep]) < RTOL
            assert rel_err(upd["critic_grad"][i], g[f"{c}/upd_grad"][ep]) < RTOL
            assert rel_err(st.critic[i], g[f"{c}/upd_params"][ep]) < RTOL
            assert rel_err(upd["adv"][i], g[f"{a}/upd_adv"][ep][..., 0]) < RTOL
            assert rel_err(upd["actor_loss"][i], g[f"{a}/upd_loss"][ep]) < RTOL
            assert rel_err(upd["actor_grad"][i], g[f"{a}/upd_grad"][ep]) < RTOL
            assert rel_err(st.actor[i], g[f"{a}/upd_params"][ep]) < RTOL
            assert np.array_equal(traj["act"][:T, :, i], g[f"{a}/upd_act"][ep].astype(np.int64))
        jt = traj["act"][:T, :, 0] * 3 + traj["act"][:T, :, 1]
        assert np.array_equal(jt, g["crit1/upd_act"][ep]) and np.array_equal(jt, g["crit2/upd_act"][ep])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_a2c_org_replay(golden, dtype):
    g = golden("a2c_org.npz")
    lr_c, lr_a, beta, gamma = g["meta_hyper"]
    T = int(g["meta_T"])
    st = L.A2COrgState(actor=g["act1/init"][0].astype(dtype), critic=g["main1/init"][0].astype(dtype),
                       lr_c=lr_c, lr_a=lr_a, beta=beta, gamma=gamma, T=T)
    sampled = list(g["act1/sampled"])
    pos = 0
    for ep in range(int(g["meta_updates"])):
        n = T + 1 if ep == 0 else T
        out = L.a2c_org_update(st, sampled[pos:pos + n])
        pos += n
        es = slice(ep * T, (ep + 1) * T)
        assert np.array_equal(out["actions"][:, 0], g["org/action"][es])
        assert np.array_equal(out["reward"][:, 0], g["org/reward"][es])          # fp64 bit-exact
        assert np.array_equal(out["states"][:, 0], g["org/obs"][es].astype(np.float32))
        assert np.array_equal(out["states"], g["main1/upd_obs"][ep])              # both rows, aliasing quirk Q4
        assert np.array_equal(out["target"], g["main1/upd_target"][ep][..., 0])   # masks == 0 -> target == reward (Q5)
        assert rel_err(out["critic_loss"], g["main1/upd_loss"][ep]) < RTOL
        assert rel_err(out["critic_grad"], g["main1/upd_grad"][ep]) < RTOL
        assert rel_err(st.critic, g["main1/upd_params"][ep]) < RTOL
        assert rel_err(out["adv"], g["act1/upd_adv"][ep][..., 0]) < RTOL
        assert rel_err(out["actor_loss"], g["act1/upd_loss"][ep]) < RTOL
        assert rel_err(out["actor_grad"], g["act1/upd_grad"][ep]) < RTOL       # includes the Q7 term through V
        assert rel_err(st.actor, g["act1/upd_params"][ep]) < RTOL
