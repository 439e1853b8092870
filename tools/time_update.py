#!/usr/bin/env python
"""ia2c_net_update at the a2c_test.py shape (65536 x 500 -> 6 -> 6 -> 6): single-pass kernel vs the kernel sequence, CUDA events,
L2 flushed between launches; and the class-API update pair (wall clock)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ia2c_b200 import _lib
from ia2c_b200.nets import ActorNetwork, CriticNetwork

lib = _lib.load()
rows, F, O = 65536, 500, 6
P = 6 * F + 6 + 36 + 6 + 6 * O + O
rng = np.random.RandomState(0)
idx = torch.from_numpy(rng.randint(0, F, size=rows))
x = torch.nn.functional.one_hot(idx, F).float().cuda()
act = torch.from_numpy(rng.randint(0, O, size=rows).astype(np.int32)).cuda()
sig = torch.randn(rows).cuda()
p = (torch.randn(P) * 0.3).cuda()
g, m, v = torch.zeros(P).cuda(), torch.zeros(P).cuda(), torch.zeros(P).cuda()
step = torch.zeros(1, dtype=torch.int32).cuda()
loss, status = torch.zeros(1).cuda(), torch.zeros(1, dtype=torch.int32).cuda()
ws = torch.empty(int(lib.ia2c_net_update_workspace(rows, F, O)), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
D = lambda t: t.data_ptr()
peak = 6535.7
for mode in ("sequence", "single"):
    if mode == "sequence":
        os.environ["IA2C_NO_SINGLE_PASS"] = "1"
    else:
        os.environ.pop("IA2C_NO_SINGLE_PASS", None)
    for kind in (0, 1):
        def call():
            _lib.check(lib.ia2c_net_update(kind, D(p), D(g), D(m), D(v), D(step), D(x), None, D(act), D(sig), 0.01, 5e-4, D(loss), D(status),
                                           D(ws), rows, F, O, _lib.stream_ptr()))
        for _ in range(3):
            call()
        tot = 0.0
        for _ in range(20):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); call(); e.record(); e.synchronize()
            tot += s.elapsed_time(e)
        us = tot / 20 * 1e3
        print(f"{mode:9s} kind {kind}: {us:7.1f} us per update, X once = {rows * F * 4 / us / 1e3:7.1f} GB/s = {rows * F * 4 / us / 1e3 / peak:.3f} of HBM peak")
os.environ.pop("IA2C_NO_SINGLE_PASS", None)
T, E = 64, 1024
for tag, obs in (("dense", x.view(T, E, F)), ("index", idx.view(T, E).cuda())):
    critic, actor = CriticNetwork("c", F, O, 5e-4), ActorNetwork("a", F, O, 1e-4, 0.01)
    a3 = act.view(T, E, 1).float()
    tg, ad = sig.view(T, E, 1), sig.view(T, E, 1)
    for _ in range(3):
        critic.batch_update(obs, a3, tg); actor.batch_update(obs, a3, ad)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        critic.batch_update(obs, a3, tg); actor.batch_update(obs, a3, ad)
    torch.cuda.synchronize()
    print(f"class API {tag}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per critic + actor update pair")
