"""World-size-2 CPU (gloo) test of the multi-rank host logic: env sharding + one all-reduce(SUM) of locally summed,
globally scaled gradients per optimiser phase reproduces the single-rank update."""
import os

import numpy as np
import pytest

from ia2c_b200.sharding import belief_draw_index, shard_envs, shard_tape


def test_shard_envs_and_tapes():
    assert shard_envs(8192, 3, 8) == (3072, 1024)
    with pytest.raises(ValueError):
        shard_envs(10, 0, 4)
    tape = np.arange(31 * 8 * 2).reshape(31, 8, 2)
    parts = [shard_tape(tape, 1, r, 4) for r in range(4)]
    assert np.array_equal(np.concatenate(parts, axis=1), tape) and shard_tape(None, 1, 0, 2) is None
    # Philox counters are global: rank 1's first belief row continues where rank 0's last one ended, and the layout is
    # the one the oracle's generator (which mirrors the kernels) uses: slots 4s..4s+3 of row r share draw r*ceil(K/4)+s
    for n_agents in (2, 3, 5, 6, 64):
        kq = (n_agents + 2) // 4
        last = belief_draw_index(0, 3, n_agents - 1, n_agents - 2, n_agents)
        first = belief_draw_index(4, 0, 0, 0, n_agents)
        assert first == (last[0] + 1, 0) and last[0] == (4 * n_agents - 1) * kq + (n_agents - 2) // 4
    from oracle import philox as P
    rows = np.array([[5 * 7 + 1]])                                  # env 5, agent 1 of N=7 (K=6: two draws)
    u = P.belief_uniforms(9, 2, 7, rows, 6)
    idx, word = belief_draw_index(4, 1, 1, 5, 7)                    # rank offset 4 + local env 1 = global env 5, slot 5
    x = P.draw(9, P.STREAM_BELIEF, 2, 7, np.array([idx]))
    assert word == 1 and u[0, 0, 5] == (float(x[1][0]) + 0.5) * 2.0 ** -32


def _worker(rank, world, port, E, out):
    import torch
    import torch.distributed as dist
    from oracle import loops as L
    from oracle import nets as NN
    from ia2c_b200.trainer import reference_init

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, M, T = 2, 5, 30
        actor, critic, fa = reference_init(N, M, seed=0)
        st = L.IA2CState(actor=actor.astype(np.float64), critic=critic.astype(np.float64), filter_action=fa)
        rng = np.random.RandomState(0)
        actions = rng.randint(0, 3, size=(T + 1, E, N))
        u = rng.rand(T + 1, E, N, N - 1)
        off, n = shard_envs(E, rank, world)
        traj = L.ia2c_rollout(st, n, actions=shard_tape(actions, 1, rank, world), u_belief=shard_tape(u, 1, rank, world))
        true_p, pred_p = L.partner_actions(traj, N)
        obs, nobs = traj["obs"][:T], traj["obs"][1:]
        rew = traj["reward"].astype(np.float32).astype(np.float64)
        flat = []
        for i in range(N):
            jt = L.joint_index(i, N, traj["act"][:T, :, i], true_p[:T, :, i])
            nja = L.joint_index(i, N, traj["act"][1:, :, i], pred_p[1:, :, i])
            loss, grad, _ = NN.critic_loss_grad(st.critic[i], obs, jt, rew, 6, 9, next_obs=nobs, next_act=nja, gamma_mask=st.gamma)
            # local SUM scaled by the GLOBAL denominator: the oracle returns local means, so rescale by n/E
            flat.append(np.concatenate([grad * (n / E), [loss * (n / E)]]))
        buf = torch.from_numpy(np.stack(flat))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        if rank == 0:
            np.save(out, buf.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from oracle import loops as L
    from oracle import nets as NN
    from ia2c_b200.trainer import reference_init

    E = 12
    out = str(tmp_path / "reduced.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, E, out), nprocs=2, join=True)
    reduced = np.load(out)
    N, M, T = 2, 5, 30
    actor, critic, fa = reference_init(N, M, seed=0)
    st = L.IA2CState(actor=actor.astype(np.float64), critic=critic.astype(np.float64), filter_action=fa)
    rng = np.random.RandomState(0)
    actions = rng.randint(0, 3, size=(T + 1, E, N))
    u = rng.rand(T + 1, E, N, N - 1)
    traj = L.ia2c_rollout(st, E, actions=actions, u_belief=u)
    true_p, pred_p = L.partner_actions(traj, N)
    rew = traj["reward"].astype(np.float32).astype(np.float64)
    for i in range(N):
        jt = L.joint_index(i, N, traj["act"][:T, :, i], true_p[:T, :, i])
        nja = L.joint_index(i, N, traj["act"][1:, :, i], pred_p[1:, :, i])
        loss, grad, _ = NN.critic_loss_grad(st.critic[i], traj["obs"][:T], jt, rew, 6, 9, next_obs=traj["obs"][1:], next_act=nja,
                                            gamma_mask=st.gamma)
        np.testing.assert_allclose(reduced[i, :147], grad, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(reduced[i, 147], loss, rtol=1e-12)
