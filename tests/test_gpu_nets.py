"""GPU parity: actor/critic MLP kernels, losses and Adam (class API through the C ABI) vs golden tapes of
the real ac_nets classes and vs the oracle.  Tolerance: 1e-5 relative (north star, fp32)."""
import numpy as np
import pytest

from oracle import nets as NN
from tests.helpers import dev, dptr, host, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.mark.parametrize("tag", ["org", "org9", "taxi", "dense"])
def test_class_api_vs_golden_updates(golden, tag):
    import torch
    import torch.nn.functional as F
    from ia2c_b200.nets import ActorNetwork, CriticNetwork

    g = golden("acnets_updates.npz")
    T, E, Fd, J, A, _ = [int(x) for x in g[f"{tag}/dims"]]
    lr_c, lr_a, beta, gamma = g[f"{tag}/hyper"]
    critic = CriticNetwork("c", Fd, J, lr_c)
    actor = ActorNetwork("a", Fd, A, lr_a, beta)
    critic.net.load_flat(g[f"{tag}/critic_init"])
    actor.net.load_flat(g[f"{tag}/actor_init"])
    assert set(critic.net.state_dict()) == {f"l{i}.{k}" for i in (1, 2, 3) for k in ("weight", "bias")}
    assert tuple(critic.net.l1.weight.shape) == (6, Fd) and tuple(actor.net.l3.weight.shape) == (A, 6)
    for it in range(3):
        k = lambda n: torch.from_numpy(g[f"{tag}/{it}/{n}"])
        obs, nobs = k("obs"), k("nobs")
        Q = critic.run_main(obs)
        Pr = actor.action_distribution(obs)
        assert Q.device.type == "cpu" and tuple(Q.shape) == (T, E, J)
        assert rel_err(Q.numpy(), k("Q").numpy()) < RTOL and rel_err(Pr.numpy(), k("P").numpy()) < RTOL
        # the script-side graph around run_main(grad=True) is plain torch on CPU tensors (ia2c.py:108-111)
        with_grad = bool(g[f"{tag}/{it}/target_has_grad"])
        qn = (critic.run_main(nobs, grad=with_grad) * F.one_hot(k("cnext"), J).float()).sum(-1, keepdims=True)
        target = k("rew").unsqueeze(-1) + gamma * k("mask").unsqueeze(-1) * qn
        assert target.requires_grad == with_grad
        assert rel_err(target.detach().numpy(), k("target").numpy()) < RTOL
        critic.batch_update(obs, k("cact"), target)
        assert rel_err(critic.losses[-1], k("critic_loss").numpy()) < RTOL
        assert rel_err(host(critic.net.flat.grad), k("critic_grad").numpy()) < RTOL
        assert rel_err(host(critic.net.flat), k("critic_params").numpy()) < RTOL
        actor.batch_update(obs, k("aact"), k("adv"))
        assert rel_err(actor.losses[-1], k("actor_loss").numpy()) < RTOL
        assert rel_err(host(actor.net.flat.grad), k("actor_grad_accum").numpy()) < RTOL   # never zeroed (Q2)
        assert rel_err(host(actor.net.flat), k("actor_params").numpy()) < RTOL
    assert np.isclose(critic.critic_loss, np.mean(critic.losses)) and len(actor.losses) == 3


@pytest.mark.parametrize("rows,F,O,softmax", [(1, 6, 3, 1), (1000, 6, 9, 0), (777, 6, 9, 1), (513, 500, 6, 0), (64, 500, 6, 1),
                                              (300, 11, 25, 1), (129, 70, 2, 0), (5000, 1, 1, 0)])
def test_mlp_forward_backward_vs_oracle(rows, F, O, softmax):
    import torch
    from ia2c_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(rows + F + O)
    flat = (rng.randn(NN.n_params(F, O)) * 0.4).astype(np.float32)
    x = rng.randn(rows, F).astype(np.float32)
    if F == 500:
        x = np.eye(F, dtype=np.float32)[rng.randint(0, F, rows)]
    dy = rng.randn(rows, O).astype(np.float32)
    y = torch.empty(rows, O, device="cuda")
    _lib.check(lib.ia2c_mlp_forward(dptr((flat)), dptr((x)), _lib.ptr(y), None, rows, F, O, 1, softmax, _lib.stream_ptr()))
    ref, cache = NN.forward(flat.astype(np.float64), x, F, O, softmax=bool(softmax), keep=True)
    assert rel_err(host(y), ref) < RTOL
    grad = torch.full((flat.size,), 7.0, device="cuda")
    dx = torch.empty(rows, F, device="cuda")
    ws = torch.empty(lib.ia2c_mlp_backward_workspace(rows, F, O), device="cuda")
    _lib.check(lib.ia2c_mlp_backward(dptr((flat)), dptr((x)), dptr((dy)), None, _lib.ptr(grad), _lib.ptr(dx), _lib.ptr(ws),
                                     rows, F, O, softmax, 0, _lib.stream_ptr()))
    dyp = dy.astype(np.float64)
    if softmax:
        dyp = ref * (dyp - (ref * dyp).sum(-1, keepdims=True))
    gref = NN.backward(flat.astype(np.float64), cache, dyp, F, O)
    assert rel_err(host(grad), gref) < RTOL
    W1 = flat[:6 * F].reshape(6, F).astype(np.float64)
    dz1 = ((dyp @ NN.unpack(flat.astype(np.float64), F, O)[4]) * (cache[4] > 0)) @ NN.unpack(flat.astype(np.float64), F, O)[2] * (cache[2] > 0)
    assert rel_err(host(dx), dz1 @ W1) < RTOL
    # accumulate mode adds onto the existing gradient; this time with the layer-1 activations saved by the forward
    h1s = torch.empty(rows, 6, device="cuda")
    _lib.check(lib.ia2c_mlp_forward(dptr((flat)), dptr((x)), _lib.ptr(y), _lib.ptr(h1s), rows, F, O, 1, softmax, _lib.stream_ptr()))
    _lib.check(lib.ia2c_mlp_backward(dptr((flat)), dptr((x)), dptr((dy)), _lib.ptr(h1s), _lib.ptr(grad), None, _lib.ptr(ws),
                                     rows, F, O, softmax, 1, _lib.stream_ptr()))
    assert rel_err(host(grad), 2 * gref) < RTOL


def test_multi_net_forward():
    import torch
    from ia2c_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(1)
    nets, rows = 5, 333
    flat = (rng.randn(nets, 105) * 0.5).astype(np.float32)
    x = rng.randn(rows, 6).astype(np.float32)
    y = torch.empty(nets, rows, 3, device="cuda")
    _lib.check(lib.ia2c_mlp_forward(dptr((flat)), dptr((x)), _lib.ptr(y), None, rows, 6, 3, nets, 1, _lib.stream_ptr()))
    for n in range(nets):
        assert rel_err(host(y)[n], NN.forward(flat[n].astype(np.float64), x, 6, 3, softmax=True)) < RTOL


def test_sampler_injected_uniforms_and_philox():
    import torch
    from ia2c_b200 import _lib
    from oracle import philox as P
    lib = _lib.load()
    rng = np.random.RandomState(2)
    rows, F, O = 20000, 6, 9
    flat = (rng.randn(NN.n_params(F, O)) * 0.7).astype(np.float32)
    x = rng.randn(rows, F).astype(np.float32)
    u = rng.rand(rows).astype(np.float32)
    act = torch.empty(rows, dtype=torch.int64, device="cuda")
    probs = torch.empty(rows, O, device="cuda")
    _lib.check(lib.ia2c_actor_sample(dptr((flat)), dptr((x)), dptr((u)), _lib.ptr(act), _lib.ptr(probs), rows, F, O, 0, 0, _lib.stream_ptr()))
    p_ref = NN.forward(flat, x, F, O, softmax=True)
    assert rel_err(host(probs), p_ref) < RTOL
    assert np.array_equal(host(act), NN.sample_inverse_cdf(host(probs), u))  # exact given the kernel's own probs
    assert (host(act) != NN.sample_inverse_cdf(p_ref, u)).mean() < 1e-3                          # oracle probs differ by ulps only
    # Philox path: the uniforms are reproducible by the oracle generator
    seed, counter = 1234567, (3 << 16) | 17
    _lib.check(lib.ia2c_actor_sample(dptr((flat)), dptr((x)), None, _lib.ptr(act), _lib.ptr(probs), rows, F, O, seed, counter, _lib.stream_ptr()))
    u2 = P.uniform_f32(seed, P.STREAM_ACTION, 3, 17, np.arange(rows))
    assert np.array_equal(host(act), NN.sample_inverse_cdf(host(probs), u2))
    freq = np.bincount(host(act), minlength=O) / rows
    assert np.abs(freq - host(probs).mean(0)).max() < 0.02


def test_q15_single_env_squeeze_is_rejected():
    import torch
    from ia2c_b200.nets import CriticNetwork
    c = CriticNetwork("c", 6, 9, 1e-3)
    obs = torch.zeros(30, 1, 6)
    with pytest.raises(ValueError, match="squeeze"):
        c.batch_update(obs, torch.zeros(30, 1), torch.zeros(30, 1, 1))
    c.batch_update(obs, torch.zeros(30, 1, 1), torch.zeros(30, 1, 1))  # explicit trailing axis is fine


def test_adam_matches_oracle_over_many_steps():
    import torch
    from ia2c_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(5)
    nets, Pn = 3, 147
    p0 = rng.randn(nets, Pn).astype(np.float32)
    p = dev(p0.copy())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    acc = torch.zeros_like(p)
    step = torch.zeros(nets, dtype=torch.int32, device="cuda")
    refs = [NN.AdamRef(Pn, 2e-4) for _ in range(nets)]
    pr = [p0[n].copy() for n in range(nets)]
    accr = np.zeros_like(p0)
    for it in range(25):
        g = (rng.randn(nets, Pn) * 10 ** rng.uniform(-4, 1)).astype(np.float32)
        _lib.check(lib.ia2c_adam_step(_lib.ptr(p), dptr((g)), _lib.ptr(acc), _lib.ptr(m), _lib.ptr(v), _lib.ptr(step),
                                      2e-4, 0.9, 0.999, 1e-8, nets, Pn, _lib.stream_ptr()))
        accr = accr + g
        for n in range(nets):
            pr[n] = refs[n].step(pr[n], accr[n])
    assert int(step[0]) == 25
    assert rel_err(host(p), np.stack(pr)) < 1e-6 and rel_err(host(acc), accr) < 1e-6


def test_index_input_path_vs_reference_tape_and_dense_path(golden):
    """integer observations stand for one_hot(idx, F) rows (a2c_test.py:57,67).  The 'taxi' tape was
    recorded from the real ac_nets classes on materialised one-hot rows; the index path must reproduce it (1e-5) and
    must equal the dense path on the same rows BIT FOR BIT — outputs, gradients, parameters, sampled actions."""
    import torch
    import torch.nn.functional as F
    from ia2c_b200.nets import ActorNetwork, CriticNetwork

    g = golden("acnets_updates.npz")
    tag = "taxi"
    T, E, Fd, J, A, _ = [int(x) for x in g[f"{tag}/dims"]]
    lr_c, lr_a, beta, gamma = g[f"{tag}/hyper"]
    torch.manual_seed(0)
    nets = {}
    for mode in ("index", "dense"):
        c, a = CriticNetwork("c", Fd, J, lr_c), ActorNetwork("a", Fd, A, lr_a, beta)
        c.net.load_flat(g[f"{tag}/critic_init"]), a.net.load_flat(g[f"{tag}/actor_init"])
        a._seed, a._calls = 1234, 0
        nets[mode] = (c, a)
    for it in range(3):
        k = lambda n: torch.from_numpy(g[f"{tag}/{it}/{n}"])
        obs, nobs = k("obs"), k("nobs")
        assert ((obs == 0) | (obs == 1)).all() and (obs.sum(-1) == 1).all()          # the tape's rows are one-hot
        views = {"index": (obs.argmax(-1), nobs.argmax(-1)), "dense": (obs, nobs)}
        out = {}
        for mode, (o, no) in views.items():
            critic, actor = nets[mode]
            Q, Pr = critic.run_main(o), actor.action_distribution(o)
            assert tuple(Q.shape) == (T, E, J) and tuple(Pr.shape) == (T, E, A)
            acts = actor.sample_action(o)
            with_grad = bool(g[f"{tag}/{it}/target_has_grad"])
            qn = (critic.run_main(no, grad=with_grad) * F.one_hot(k("cnext"), J).float()).sum(-1, keepdims=True)
            target = k("rew").unsqueeze(-1) + gamma * k("mask").unsqueeze(-1) * qn
            critic.batch_update(o, k("cact"), target)
            actor.batch_update(o, k("aact"), k("adv"))
            out[mode] = dict(Q=Q.numpy(), P=Pr.numpy(), acts=acts.numpy(), probs=host(actor.last_probs),
                             cgrad=host(critic.net.flat.grad), cpar=host(critic.net.flat), agrad=host(actor.net.flat.grad),
                             apar=host(actor.net.flat), closs=critic.losses[-1], aloss=actor.losses[-1])
        for name in out["index"]:
            assert np.array_equal(out["index"][name], out["dense"][name]), (name, it)
        o = out["index"]
        assert rel_err(o["Q"], k("Q").numpy()) < RTOL and rel_err(o["P"], k("P").numpy()) < RTOL
        assert rel_err(o["closs"], k("critic_loss").numpy()) < RTOL and rel_err(o["aloss"], k("actor_loss").numpy()) < RTOL
        assert rel_err(o["cgrad"], k("critic_grad").numpy()) < RTOL and rel_err(o["cpar"], k("critic_params").numpy()) < RTOL
        assert rel_err(o["agrad"], k("actor_grad_accum").numpy()) < RTOL and rel_err(o["apar"], k("actor_params").numpy()) < RTOL


@pytest.mark.parametrize("rows,F,O,softmax", [(1, 500, 6, 0), (4097, 500, 6, 1), (65536, 500, 6, 0), (300, 6, 9, 0), (1000, 3000, 4, 1)])
def test_index_kernels_equal_dense_kernels(rows, F, O, softmax):
    import torch
    from ia2c_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(rows + F)
    P = 6 * F + 6 + 36 + 6 + 6 * O + O
    params = torch.from_numpy((rng.randn(P) * 0.3).astype(np.float32)).cuda()
    idx = torch.from_numpy(rng.randint(0, F, size=rows)).cuda()
    if rows > 2:
        idx[0], idx[1] = 0, F - 1
    x = torch.nn.functional.one_hot(idx, F).float().contiguous()
    dy = torch.from_numpy(rng.randn(rows, O).astype(np.float32)).cuda()
    st = _lib.stream_ptr()
    res = {}
    for mode in ("dense", "index"):
        y, h1 = torch.empty(rows, O, device="cuda"), torch.empty(rows, 6, device="cuda")
        grad = torch.empty(P, device="cuda")
        ws = torch.empty(lib.ia2c_mlp_backward_workspace(rows, F, O), device="cuda")
        if mode == "dense":
            _lib.check(lib.ia2c_mlp_forward(_lib.ptr(params), _lib.ptr(x), _lib.ptr(y), _lib.ptr(h1), rows, F, O, 1, softmax, st))
            _lib.check(lib.ia2c_mlp_backward(_lib.ptr(params), _lib.ptr(x), _lib.ptr(dy), _lib.ptr(h1), _lib.ptr(grad), None,
                                             _lib.ptr(ws), rows, F, O, softmax, 0, st))
        else:
            _lib.check(lib.ia2c_mlp_forward_index(_lib.ptr(params), _lib.ptr(idx), _lib.ptr(y), _lib.ptr(h1), rows, F, O, softmax, st))
            _lib.check(lib.ia2c_mlp_backward_index(_lib.ptr(params), _lib.ptr(idx), _lib.ptr(dy), _lib.ptr(h1), _lib.ptr(grad),
                                                   _lib.ptr(ws), rows, F, O, softmax, 0, st))
        res[mode] = (host(y), host(h1), host(grad))
    for a, b, name in zip(res["dense"], res["index"], ("y", "h1", "grad")):
        assert np.array_equal(a, b), name


def test_index_input_rejects_out_of_range_classes():
    import torch
    from ia2c_b200.nets import CriticNetwork
    c = CriticNetwork("c", 500, 6, 1e-3)
    with pytest.raises(RuntimeError, match="Class values"):
        c.run_main(torch.tensor([[3, 500]]))
    # batch_update validates on the device (no extra reduction + sync): the offending rows are clamped, the call still raises
    from ia2c_b200.nets import ActorNetwork
    a = ActorNetwork("a", 500, 6, 1e-3, 0.01)
    obs = torch.tensor([[3, 499], [0, 7]]).cuda()
    act, sig = torch.zeros(2, 2, 1), torch.ones(2, 2, 1)
    c.batch_update(obs, act, sig)
    a.batch_update(obs, act, sig)
    bad = obs.clone()
    bad[1, 1] = 500
    with pytest.raises(RuntimeError, match="Class values"):
        c.batch_update(bad, act, sig)
    with pytest.raises(RuntimeError, match="Class values"):
        a.batch_update(bad.cpu() - 501, act, sig)


@pytest.mark.parametrize("kind,rows,F,O", [(0, 5000, 500, 6), (1, 5000, 500, 6), (0, 1024 + 17, 64, 8), (1, 2048, 128, 3), (1, 4100, 500, 6)])
def test_single_pass_update_equals_the_kernel_sequence(kind, rows, F, O):
    """ia2c_net_update: the single-pass kernel for wide dense rows (bulk async copies, X read once) against the forward / loss /
    backward / Adam kernel sequence behind the same entry point (IA2C_NO_SINGLE_PASS), which the golden tapes pin."""
    import os
    import torch
    from ia2c_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(rows + F + O + kind)
    P = NN.n_params(F, O)
    flat0 = (rng.randn(P) * 0.3).astype(np.float32)
    x = np.zeros((rows, F), dtype=np.float32)
    x[np.arange(rows), rng.randint(0, F, size=rows)] = 1.0          # one-hot rows (a2c_test.py:57) ...
    x[: rows // 2] += (rng.rand(rows // 2, F) < 0.05) * rng.randn(rows // 2, F).astype(np.float32)   # ... and some dense ones
    act = rng.randint(0, O, size=rows).astype(np.int32)
    sig = rng.randn(rows).astype(np.float32)
    grad0 = (rng.randn(P) * 0.01).astype(np.float32)               # the actor's running gradient sum is not zero
    out = {}
    for mode in ("sequence", "single"):
        if mode == "sequence":
            os.environ["IA2C_NO_SINGLE_PASS"] = "1"
        else:
            os.environ.pop("IA2C_NO_SINGLE_PASS", None)
        p, g = dev(flat0.copy()), dev(grad0.copy())
        m, v = dev(np.zeros(P, np.float32)), dev(np.zeros(P, np.float32))
        step = torch.zeros(1, dtype=torch.int32, device="cuda")
        loss, status = torch.zeros(1, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
        ws = torch.empty(int(lib.ia2c_net_update_workspace(rows, F, O)), dtype=torch.float32, device="cuda")
        xd, ad, sd = dev(x), dev(act), dev(sig)
        for it in range(2):
            P_ = lambda t: t.data_ptr()
            _lib.check(lib.ia2c_net_update(kind, P_(p), P_(g), P_(m), P_(v), P_(step), P_(xd), None, P_(ad), P_(sd), 0.01,
                                           5e-4, P_(loss), P_(status), P_(ws), rows, F, O, None))
        torch.cuda.synchronize()
        out[mode] = (host(p), host(g), float(loss.item()), int(step.item()), int(status.item()))
    os.environ.pop("IA2C_NO_SINGLE_PASS", None)
    (p1, g1, l1, s1, st1), (p2, g2, l2, s2, st2) = out["sequence"], out["single"]
    assert s1 == s2 == 2 and st1 == st2 == 0
    assert rel_err(l2, l1) < RTOL and rel_err(g2, g1) < RTOL and rel_err(p2, p1) < RTOL
    assert not np.array_equal(p1, flat0)
