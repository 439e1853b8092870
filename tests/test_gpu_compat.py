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
