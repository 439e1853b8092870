"""The whole-episode belief kernel at the config-5 per-GPU shape (1024 envs x 256 agents, 31 steps) — timing and the ncu target."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ia2c_b200 import _lib  # noqa: E402

lib = _lib.load()
E, N, M, T1 = int(os.environ.get("EP_ENVS", 1024)), int(os.environ.get("EP_AGENTS", 256)), 5, 31
K = N - 1
rng = np.random.RandomState(0)
fa = rng.rand(N, M, 3)
fa /= fa.sum(-1, keepdims=True)
fa_d = torch.from_numpy(fa).cuda()
rec = torch.zeros(E, N, K, 8, dtype=torch.uint8, device="cuda")
partner = torch.empty(T1, E, N, dtype=torch.uint8, device="cuda")
act = torch.randint(0, 3, (T1, E, N), dtype=torch.uint8, device="cuda")


def run(ep):
    _lib.check(lib.ia2c_belief_update_pairs_episode(_lib.ptr(rec), _lib.ptr(fa_d), _lib.ptr(act), None, None, None, _lib.ptr(partner), E, N, M, T1, 7, ep,
                                                    0, _lib.stream_ptr()))


for ep in range(2):
    run(ep)
best = 1e9
for g in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(2 + g)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
upd = E * N * K * T1
print(f"episode kernel: {best:.3f} ms per episode = {best / T1 * 1e3:.1f} us per step-equivalent, {upd / best / 1e6:.1f} G updates/s")
