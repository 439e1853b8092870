"""Oracle pin: oracle/nets.py closed forms against values / autograd gradients / Adam steps recorded from
the real ac_nets classes (tests/golden/acnets_updates.npz)."""
import numpy as np
import pytest

from oracle import nets as NN

RTOL = 1e-5  # north-star tolerance (fp32, relative to the tensor's scale)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("tag", ["org", "org9", "taxi", "dense"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_acnets_updates(golden, tag, dtype):
    g = golden("acnets_updates.npz")
    T, E, F, J, A, _ = [int(x) for x in g[f"{tag}/dims"]]
    lr_c, lr_a, beta, gamma = g[f"{tag}/hyper"]
    critic = g[f"{tag}/critic_init"].astype(dtype)
    actor = g[f"{tag}/actor_init"].astype(dtype)
    assert critic.size == NN.n_params(F, J) and actor.size == NN.n_params(F, A)
    adam_c, adam_a = NN.AdamRef(critic.size, lr_c, dtype), NN.AdamRef(actor.size, lr_a, dtype)
    accum = np.zeros_like(actor)
    for it in range(3):
        k = lambda n: g[f"{tag}/{it}/{n}"]
        assert rel_err(NN.forward(critic, k("obs"), F, J).reshape(T, E, J), k("Q")) < RTOL
        assert rel_err(NN.forward(actor, k("obs"), F, A, softmax=True).reshape(T, E, A), k("P")) < RTOL
        if bool(k("target_has_grad")):
            loss, grad, target = NN.critic_loss_grad(critic, k("obs"), k("cact"), k("rew"), F, J, next_obs=k("nobs"),
                                                     next_act=k("cnext"), gamma_mask=gamma * k("mask"))
            assert rel_err(target.reshape(T, E, 1), k("target")) < RTOL
        else:
            loss, grad, _ = NN.critic_loss_grad(critic, k("obs"), k("cact"), k("target"), F, J)
        assert rel_err(loss, k("critic_loss")) < RTOL
        assert rel_err(grad, k("critic_grad")) < RTOL
        critic = adam_c.step(critic, grad)
        assert rel_err(critic, k("critic_params")) < RTOL
        loss, grad, _ = NN.actor_loss_grad(actor, k("obs"), k("aact"), k("adv"), beta, F, A)
        accum = accum + grad
        assert rel_err(loss, k("actor_loss")) < RTOL
        assert rel_err(accum, k("actor_grad_accum")) < RTOL  # running sum, never zeroed (Q2)
        actor = adam_a.step(actor, accum)
        assert rel_err(actor, k("actor_params")) < RTOL


def test_clamp_edge_cases():
    # q_a > 1 - eps: log-prob term is log(1-eps) with zero gradient
    F, A = 6, 3
    flat = np.zeros(NN.n_params(F, A), dtype=np.float64)
    flat[-3:] = [40.0, 0.0, 0.0]
    loss, grad, dl = NN.actor_loss_grad(flat, np.ones((1, F)), [0], [2.0], 0.0, F, A)
    assert np.isclose(loss, -2.0 * np.log(1 - NN.EPS_CLAMP))
    assert np.abs(grad).max() == 0.0
