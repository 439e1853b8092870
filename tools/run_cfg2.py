"""A few cfg2 episodes (4096 envs x 2 agents x 30 steps, fused rollout) — the command ncu captures are taken on."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ia2c_b200.trainer import IA2CTrainer, reference_init  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tr = IA2CTrainer(4096, n_agents=2, init=reference_init(2, 5, seed=0), seed=7, fused_rollout=True)
for _ in range(n):
    tr.train_episode()
torch.cuda.synchronize()
print("ok", tr.read_stats()["mean_return"])
