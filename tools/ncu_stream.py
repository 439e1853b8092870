"""The HBM-streaming kernels at bench.py's `kernels` sizes (16.8 M env-steps at N=2, 2.1 M env-steps at N=256, 8.4 M dense belief updates) — the command
their ncu captures are taken on; prints the event-timed GB/s of the same launches (algorithmic bytes: 54 B, 308 B and 88 B per unit)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ia2c_b200 import _lib  # noqa: E402
from ia2c_b200.org_env import OrgVecEnv  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
st = _lib.stream_ptr()
E = 1 << 24
env = OrgVecEnv(E, n_agents=2)
act = torch.randint(0, 3, (E, 2), dtype=torch.uint8, device=dev)
Ew, Nw = 1 << 21, 256
envw = OrgVecEnv(Ew, n_agents=Nw)
actw = torch.randint(0, 3, (Ew, Nw), dtype=torch.uint8, device=dev)
R = 1 << 23
fa = torch.rand(5, 3, dtype=torch.float64, device=dev)
fa /= fa.sum(1, keepdim=True)
lik = torch.full((R, 3), 0.1, dtype=torch.float64, device=dev)
lik[torch.arange(R, device=dev), torch.randint(0, 3, (R,), device=dev)] = 0.8
prev = torch.full((R, 5), 0.2, dtype=torch.float64, device=dev)
u = torch.rand(R, dtype=torch.float64, device=dev)
ap = torch.empty(R, dtype=torch.int64, device=dev)
bp = torch.empty(R, 5, dtype=torch.float64, device=dev)


def env_step():
    _lib.check(lib.ia2c_org_step_agents(_lib.ptr(env.state), _lib.ptr(env.hist), _lib.ptr(env.cls), None, _lib.ptr(act),
                                        _lib.ptr(env.obs), None, _lib.ptr(env.reward_f32), None, None, E, 2, 0, st))


def env_step_warp():
    _lib.check(lib.ia2c_org_step_agents(_lib.ptr(envw.state), _lib.ptr(envw.hist), _lib.ptr(envw.cls), None, _lib.ptr(actw),
                                        _lib.ptr(envw.obs), None, _lib.ptr(envw.reward_f32), None, None, Ew, Nw, 0, st))


def dense():
    _lib.check(lib.ia2c_belief_update_dense(_lib.ptr(fa), _lib.ptr(lik), _lib.ptr(prev), _lib.ptr(u), _lib.ptr(ap), _lib.ptr(bp),
                                            None, R, 5, 3, st))


for name, fn, nbytes in (("org_step_thread_kernel", env_step, 54 * E), ("org_step_warp32_kernel", env_step_warp, (52 + Nw) * Ew),
                         ("belief_dense_kernel<5,3>", dense, 88 * R)):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record()
    b.synchronize()
    sec = a.elapsed_time(b) / 5 * 1e-3
    print(f"{name}: {sec * 1e6:.1f} us per launch, {nbytes / sec / 1e9:.0f} GB/s algorithmic")
