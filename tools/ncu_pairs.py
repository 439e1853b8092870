"""A few pairwise belief updates at the config-5 per-GPU shape (1024 envs x 256 agents) — the command the ncu
capture of belief_pairs_table_kernel is taken on."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ia2c_b200 import _lib  # noqa: E402

lib = _lib.load()
E, N, M = 1024, 256, 5
K = N - 1
rng = np.random.RandomState(0)
fa = rng.rand(N, M, 3)
fa /= fa.sum(-1, keepdims=True)
fa_d = torch.from_numpy(fa).cuda()
rec = torch.zeros(E, N, K, 8, dtype=torch.uint8, device="cuda")
partner = torch.empty(E, N, dtype=torch.uint8, device="cuda")
pred = torch.empty(E, N, K, dtype=torch.uint8, device="cuda")
acts = [torch.randint(0, 3, (E, N), dtype=torch.uint8, device="cuda") for _ in range(4)]


def update(t, with_pred):
    _lib.check(lib.ia2c_belief_update_pairs(_lib.ptr(rec), _lib.ptr(fa_d), _lib.ptr(acts[t % 4]), None,
                                            _lib.ptr(pred) if with_pred else None, None, _lib.ptr(partner),
                                            E, N, M, int(t == 0), 7, 0, t, 0, _lib.stream_ptr()))


for t in range(4):
    update(t, False)
for with_pred in (False, True, False, True):
    best = 1e9
    for g in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for t in range(4, 14):
            update(t, with_pred)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 10)
    print("pred_out" if with_pred else "fast", "ms per update", round(best, 4), "records/s", E * N * K / (best * 1e-3))
