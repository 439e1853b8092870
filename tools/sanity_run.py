"""Touches every kernel of libia2c_b200.so once at small, ragged sizes (for compute-sanitizer runs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ia2c_b200.belief import BeliefFilter
from ia2c_b200.nets import ActorNetwork, CriticNetwork
from ia2c_b200.org_env import Org, OrgVecEnv
from ia2c_b200.trainer import IA2CTrainer, reference_init

rng = np.random.RandomState(0)
for E, N in ((37, 2), (5, 3), (3, 9), (2, 33)):
    env = OrgVecEnv(E, n_agents=N, max_episode_steps=4)
    for _ in range(6):
        env.step(rng.randint(0, 3, size=(E, N)).astype(np.uint8))
    if N == 2:
        env.step(rng.randint(-1, 10, size=E))
single = Org()
single.reset()
for a in (0, 8, 4, 11):
    single.step(a)
bf = BeliefFilter(5, 3, 19)
prior = bf.prior
for _ in range(3):
    lik = np.where(rng.rand(19, 3) < 0.34, 0.8, 0.1)
    _, prior, _ = bf.update(lik, prior)
BeliefFilter(3, 4, 7).update(rng.rand(7, 4), np.full((7, 3), 0.33))
for F_, O in ((6, 9), (500, 6), (11, 25)):
    c, a = CriticNetwork("c", F_, O, 1e-3), ActorNetwork("a", F_, O, 1e-3, 0.01)
    obs = torch.randn(7, 5, F_)
    act = torch.randint(0, O, (7, 5, 1)).float()
    a.sample_action(obs)
    tgt = torch.randn(7, 5, 1) + 0.9 * c.run_main(obs, grad=True).sum(-1, keepdim=True)
    c.batch_update(obs, act, tgt)
    a.batch_update(obs, act, torch.randn(7, 5, 1))
for E, N, fused in ((37, 2, True), (37, 2, False), (9, 5, True), (3, 33, False), (2, 64, False)):
    tr = IA2CTrainer(E, n_agents=N, init=reference_init(N, 5, seed=1), seed=3, fused_rollout=fused, dumps=True, steps_per_episode=7,
                     max_episode_steps=7)
    tr.train_episode()
    T = 7
    tr.inject(u_action=rng.rand(T + 1, E, N).astype(np.float32), u_belief=rng.rand(T + 1, E, N, N - 1))
    tr.train_episode(sync_stats=True)
    tapes = [tr.pack_host_tape(rng.rand(T + 1, E, N), rng.rand(T + 1, E, N, N - 1)) for _ in range(3)]
    tr.train_episodes_host(tapes)
    tr.inject(actions=rng.randint(0, 3, size=(T + 1, E, N)), u_belief=rng.rand(T + 1, E, N, N - 1))
    tr.train_episode(sync_stats=True)
    tr.train_episode_timed()
torch.cuda.synchronize()
print("SANITY_RUN_OK")
