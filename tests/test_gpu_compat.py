"""GPU parity of the drop-in class API on the a2c_org_test.py flow (config 1): a driver written against the
compat modules (Org, ac_nets) replays the reference's sampled actions and must reproduce the reference run
recorded in tests/golden/a2c_org.npz, including quirks Q4-Q7."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import ROOT
from tests.helpers import host, rel_err

pytestmark = pytest.mark.gpu


def test_a2c_org_flow_through_compat_modules(golden):
    g = golden("a2c_org.npz")
    sys.path.insert(0, os.path.join(ROOT, "ia2c_b200", "compat"))
    try:
        import torch
        from ac_nets import ActorNetwork, CriticNetwork, F
        from Org import Org
    finally:
        sys.path.pop(0)
    critic_lr, actor_lr, ent_coef, gamma = g["meta_hyper"]
    n_steps, n_rows, n_obs, n_act = int(g["meta_T"]), 2, 6, 9
    env = Org()
    critic = CriticNetwork("main1", n_obs, n_act, critic_lr)
    actor = ActorNetwork("act1", n_obs, n_act, actor_lr, ent_coef)
    critic.net.load_flat(g["main1/init"][0])
    actor.net.load_flat(g["act1/init"][0])
    actor.replay(g["act1/sampled"])               # the reference's torch.multinomial draws, injected
    states, _ = env.reset(seed=42)
    actions = actor.sample_action(torch.tensor(states).float(), grad=True)
    for upd in range(int(g["meta_updates"])):
        buf_s = torch.zeros(n_steps, n_rows, n_obs)
        buf_ns = torch.zeros(n_steps, n_rows, n_obs)
        buf_a = torch.zeros(n_steps, n_rows, 1)
        buf_na = torch.zeros(n_steps, n_rows, 1)
        buf_r = torch.zeros(n_steps, n_rows)
        masks = torch.zeros(n_steps, n_rows)
        for t in range(n_steps):
            nxt, rew, terminated, truncated, _ = env.step(actions.detach().cpu().numpy())
            nxt_actions = actor.sample_action(torch.tensor(nxt).float(), grad=True)
            buf_r[t] = torch.tensor(rew)
            buf_s[t] = torch.tensor(states)     # same array object as nxt after the first step (Q4)
            buf_ns[t] = torch.tensor(nxt)
            buf_a[t] = actions.unsqueeze(-1)
            buf_na[t] = nxt_actions.unsqueeze(-1)
            masks[t] = torch.tensor(terminated)  # always False -> 0 (Q5)
            assert rew == g["org/reward"][upd * n_steps + t] and env.state == g["org/state"][upd * n_steps + t]
            states, actions = nxt, nxt_actions
        q_next = critic.run_main(buf_ns, grad=True)
        oh_next = F.one_hot(buf_na.squeeze(-1).long(), num_classes=n_act).detach().float()
        target = buf_r.unsqueeze(-1) + gamma * masks.unsqueeze(-1) * (q_next * oh_next).sum(-1, keepdims=True)
        assert np.array_equal(buf_s.numpy(), g["main1/upd_obs"][upd]) and torch.equal(buf_s, buf_ns)
        assert np.array_equal(target.detach().numpy(), g["main1/upd_target"][upd])
        critic.batch_update(buf_s, buf_a, target)
        assert rel_err(critic.losses[-1], g["main1/upd_loss"][upd]) < 1e-5
        assert rel_err(host(critic.net.flat.grad), g["main1/upd_grad"][upd]) < 1e-5
        assert rel_err(host(critic.net.flat), g["main1/upd_params"][upd]) < 1e-5
        Q = critic.run_main(buf_s, grad=False)
        dist = actor.action_distribution(buf_s, grad=True)
        V = (Q * dist).sum(-1, keepdims=True)
        oh = F.one_hot(buf_a.squeeze(-1).long(), num_classes=n_act).float()
        adv = (Q * oh).sum(-1, keepdims=True) - V
        assert adv.requires_grad                                   # differentiable advantage (Q7)
        assert rel_err(adv.detach().numpy(), g["act1/upd_adv"][upd]) < 1e-5
        actor.batch_update(buf_s, buf_a, adv)
        assert rel_err(actor.losses[-1], g["act1/upd_loss"][upd]) < 1e-5
        assert rel_err(host(actor.net.flat.grad), g["act1/upd_grad"][upd]) < 1e-5
        assert rel_err(host(actor.net.flat), g["act1/upd_params"][upd]) < 1e-5
    assert rel_err(critic.critic_loss, g["critic_loss_window"]) < 1e-5


def test_ia2c_flow_through_compat_modules_and_vector_env(golden):
    """The ia2c.py flow (config 0/2 shape: 2 agents, E vector envs, belief filters) written against the drop-in modules
    and the gymnasium-compatible GPU vector env, replaying the reference's recorded samples and np.random draws."""
    g = golden("ia2c_E10.npz")
    for sub in ("compat", "compat_gym"):
        sys.path.insert(0, os.path.join(ROOT, "ia2c_b200", sub))
    try:
        import gymnasium as gym
        import torch
        from gymnasium.envs.registration import register
        from ac_nets import ActorNetwork, CriticNetwork, F
        from belief_filter import BeliefFilter
    finally:
        del sys.path[:2]
    lr_c, lr_a, beta, gamma = g["meta_hyper"]
    E, T = int(g["meta_n_envs"]), int(g["meta_T"])
    register(id="Org-v0", entry_point="Org:Org", max_episode_steps=30)
    envs = gym.make_vec("Org-v0", num_envs=E)
    critics = [CriticNetwork(f"crit{k + 1}", 6, 9, lr_c) for k in range(2)]
    actors = [ActorNetwork(f"act{k + 1}", 6, 3, lr_a, beta) for k in range(2)]
    for k in range(2):
        critics[k].net.load_flat(g[f"crit{k + 1}/init"][0])
        actors[k].net.load_flat(g[f"act{k + 1}/init"][0])
        actors[k].replay(g[f"act{k + 1}/sampled"])
    np.random.seed(0)
    bfs = [BeliefFilter(5, 3, E), BeliefFilter(5, 3, E)]
    for k in range(2):
        bfs[k].filterAction = g[f"bf{k}/filterAction"][0].copy()
        bfs[k].filters = bfs[k].filterAction.transpose()
    u_tape = {0: iter(g["bf0/u"]), 1: iter(g["bf1/u"])}
    real_rand = np.random.rand

    def likelihood(other):
        p = np.ones((E, 3)) * 0.1
        p[np.arange(E), np.asarray(other)] = 0.8
        return p

    def belief(k, other, prior):
        np.random.rand = lambda *shape: next(u_tape[k])
        try:
            return bfs[k].update(likelihood(other), prior)
        finally:
            np.random.rand = real_rand

    step_idx, call_idx = 0, 0
    for ep in range(int(g["meta_episodes"])):
        obs = torch.zeros(T, E, 6); nobs = torch.zeros(T, E, 6); rew = torch.zeros(T, E)
        ta = torch.zeros(T, E, 2); pa = torch.zeros(T, E, 2); tna = torch.zeros(T, E, 2); pna = torch.zeros(T, E, 2)
        s, _ = envs.reset()
        assert np.array_equal(s, g["env/reset_obs"][ep])
        a = [actors[k].sample_action(torch.tensor(s), grad=True) for k in range(2)]
        pred = [None, None]
        priors = [bfs[0].prior, bfs[1].prior]
        pred[1], priors[0], _ = belief(0, a[1], priors[0])   # bf1 predicts agent 2 from agent 1's private observation
        pred[0], priors[1], _ = belief(1, a[0], priors[1])
        for k in range(2):
            assert np.array_equal(priors[k], g[f"bf{k}/bprime"][call_idx])
        call_idx += 1
        for t in range(T):
            joint = a[0] * 3 + a[1] % 3
            s_, r, done, term, info = envs.step(joint.detach())
            assert np.array_equal(s_, g["env/obs"][step_idx]) and np.array_equal(r, g["env/reward"][step_idx])
            assert s_.dtype == np.float32 and r.dtype == np.float64
            a_ = [actors[k].sample_action(torch.tensor(s_), grad=True) for k in range(2)]
            pred_ = [None, None]
            pred_[1], priors[0], _ = belief(0, a_[1], priors[0])
            pred_[0], priors[1], _ = belief(1, a_[0], priors[1])
            for k in range(2):
                assert np.array_equal(priors[k], g[f"bf{k}/bprime"][call_idx]) and np.array_equal(pred_[1 - k], g[f"bf{k}/ap"][call_idx])
            obs[t], nobs[t], rew[t] = torch.tensor(s), torch.tensor(s_), torch.tensor(r)
            ta[t] = torch.stack([a[0], a[1]], 1); tna[t] = torch.stack([a_[0], a_[1]], 1)
            pa[t] = torch.tensor(np.stack([pred[0], pred[1]], 1)); pna[t] = torch.tensor(np.stack([pred_[0], pred_[1]], 1))
            s, a, pred = s_, a_, pred_
            step_idx += 1
            call_idx += 1
        nja = [tna[:, :, 0].int() * 3 + pna[:, :, 1].int() % 3, pna[:, :, 0].int() * 3 + tna[:, :, 1].int() % 3]
        jt = ta[:, :, 0].int() * 3 + ta[:, :, 1].int() % 3
        ja = [ta[:, :, 0].int() * 3 + pa[:, :, 1].int() % 3, pa[:, :, 0].int() * 3 + ta[:, :, 1].int() % 3]
        oh_next = [F.one_hot(x.long(), 9).float() for x in nja]
        for k in range(2):
            qn = (critics[k].run_main(nobs, grad=True) * oh_next[k]).sum(-1, keepdims=True)
            critics[k].batch_update(obs, jt, rew.unsqueeze(-1) + gamma * qn)     # residual-gradient target (Q8)
            name = f"crit{k + 1}"
            assert rel_err(critics[k].losses[-1], g[f"{name}/upd_loss"][ep]) < 1e-5
            assert rel_err(host(critics[k].net.flat), g[f"{name}/upd_params"][ep]) < 1e-5
        for k in range(2):
            qn = (critics[k].run_main(nobs) * oh_next[k]).sum(-1, keepdims=True)
            qc = (critics[k].run_main(obs) * F.one_hot(ja[k].long(), 9).float()).sum(-1, keepdims=True)
            adv = rew.unsqueeze(-1) + gamma * qn - qc
            actors[k].batch_update(obs, ta[:, :, k], adv)
            name = f"act{k + 1}"
            assert rel_err(adv.numpy(), g[f"{name}/upd_adv"][ep]) < 1e-5
            assert rel_err(host(actors[k].net.flat), g[f"{name}/upd_params"][ep]) < 1e-5
