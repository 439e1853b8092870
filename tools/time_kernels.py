"""Warm per-kernel durations (CUDA events between the kernels, ia2c_train_episode_timed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ia2c_b200.trainer import IA2CTrainer, reference_init  # noqa: E402

configs = [(4096, 2, True, True, "pipe"), (4096, 2, True, True, "columns"), (4096, 2, True, False, "pipe"),
           (4096, 2, False, False, "pipe"), (1024, 64, False, False, "pipe"), (1024, 64, False, False, "columns"),
           (1024, 256, False, False, "pipe"), (1024, 256, False, False, "columns")]
for (E, N, fused, fc, ak) in configs:
    tr = IA2CTrainer(E, n_agents=N, init=reference_init(N, 5, seed=0), seed=7, fused_rollout=fused, fused_critic=fc,
                     actor_kernel=ak)
    for _ in range(5):
        tr.train_episode()
    acc, n = {}, 30
    for _ in range(n):
        for k, v in tr.train_episode_timed().items():
            acc[k] = acc.get(k, 0) + v / n
    import torch
    for _once in (0,):
        t2 = IA2CTrainer(E, n_agents=N, init=reference_init(N, 5, seed=0), seed=7, fused_rollout=fused, fused_critic=fc,
                         actor_kernel=ak)
        for _ in range(5):
            t2.train_episode()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(50):
            t2.train_episode()
        b.record()
        torch.cuda.synchronize()
        print(f"   back-to-back episodes: {a.elapsed_time(b) / 50 * 1000:.1f} us")
    print(f"E={E} N={N} fused_rollout={fused} fused_critic={fc} actor={ak}", {k: round(v * 1000, 2) for k, v in acc.items()}, "us",
          "total", round(sum(acc.values()) * 1000, 1))
