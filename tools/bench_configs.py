#!/usr/bin/env python
"""Per-config measurements for BASELINE.json configs 2-5 (bench.py covers config 2 as the headline).

    python tools/bench_configs.py [--configs cfg3,cfg4,cfg5] [--steps 5]

Prints one JSON line per config.  cfg4/cfg5 are Org-N (builder-defined many-agent extension);
cfg5 is the per-GPU share of the 8-GPU configuration (8192 envs / 8 = 1024 envs x 256 agents).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch


def ev():
    return torch.cuda.Event(enable_timing=True)


def time_trainer(name, E, N, steps, warmup=2, fused=False):
    from ia2c_b200 import _lib
    from ia2c_b200.trainer import IA2CTrainer, reference_init
    tr = IA2CTrainer(E, n_agents=N, init=reference_init(N, 5, seed=0), seed=7, fused_rollout=fused)
    for _ in range(warmup):
        tr.train_episode()
    torch.cuda.synchronize()
    marks = [[ev() for _ in range(4)] for _ in range(steps)]
    l0 = _lib.launch_count()
    for m in marks:
        m[0].record(); tr.rollout(); m[1].record()
        d, s = tr.desc, tr._stream()
        import ctypes as C
        _lib.check(tr.lib.ia2c_critic_phase(C.byref(d), s)); m[2].record()
        _lib.check(tr.lib.ia2c_actor_phase(C.byref(d), s)); m[3].record()
        tr.episode += 1
    torch.cuda.synchronize()
    launches = (_lib.launch_count() - l0) / steps
    seg = lambda a, b: sum(m[a].elapsed_time(m[b]) for m in marks) / steps
    total = seg(0, 3)
    units = E * N * 30
    pairs = E * N * (N - 1) * 31
    out = {"config": name, "label": "Org-N (builder-defined)" if N > 2 else "reference 2-agent Org", "envs": E, "agents": N,
           "ms_per_episode": total, "agent_steps_per_s": units / (total * 1e-3), "rollout_ms": seg(0, 1), "critic_ms": seg(1, 2),
           "actor_ms": seg(2, 3), "belief_pair_updates_per_s_in_rollout": pairs / (seg(0, 1) * 1e-3), "launches_per_episode": launches,
           "fused_rollout": fused}
    del tr
    torch.cuda.empty_cache()
    return out


def time_acnets(steps):
    """cfg3: ac_nets critic + actor batch_update at the a2c_test.py shape (500 -> 6 -> 6 -> 6), batch 65536."""
    from ia2c_b200.nets import ActorNetwork, CriticNetwork
    T, E, F, O = 64, 1024, 500, 6
    rng = np.random.RandomState(0)
    out = []
    for tag, Fd, Od in (("taxi 500->6->6->6", 500, 6), ("taxi 500->6->6->6, index input (no one-hot tensor)", 500, 6),
                        ("org 6->6->6->9", 6, 9)):
        idx = torch.from_numpy(rng.randint(0, Fd, size=(T, E)))
        if "index" in tag:
            obs = idx.cuda()
        else:
            obs = (torch.nn.functional.one_hot(idx, Fd).float() if Fd == 500 else torch.randn(T, E, Fd)).cuda()
        act = torch.from_numpy(rng.randint(0, Od, size=(T, E, 1)).astype(np.float32)).cuda()
        target = torch.randn(T, E, 1).cuda()
        adv = torch.randn(T, E, 1).cuda()
        critic = CriticNetwork("c", Fd, Od, 5e-4)
        actor = ActorNetwork("a", Fd, Od, 1e-4, 0.01)
        for _ in range(2):
            critic.batch_update(obs, act, target); actor.batch_update(obs, act, adv)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            critic.batch_update(obs, act, target); actor.batch_update(obs, act, adv)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        x_bytes = T * E * Fd * 4
        out.append({"config": "cfg3", "shape": tag, "batch": T * E, "ms_per_update_pair": dt * 1e3, "rows_per_s": T * E / dt,
                    "x_gbs_if_read_4x": 4 * x_bytes / dt / 1e9,
                    "note": "critic+actor batch_update through the class API (fwd, loss, bwd, Adam; host sync for the loss window)"})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg2,cfg3,cfg4,cfg5")
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    want = a.configs.split(",")
    if "cfg2" in want:
        print(json.dumps(time_trainer("cfg2", 4096, 2, max(a.steps, 50), fused=True)))
    if "cfg3" in want:
        for r in time_acnets(max(a.steps, 10)):
            print(json.dumps(r))
    if "cfg4" in want:
        print(json.dumps(time_trainer("cfg4", 1024, 64, a.steps)))
    if "cfg5" in want:
        print(json.dumps(time_trainer("cfg5 (per-GPU share: 8192/8 envs)", 1024, 256, a.steps)))


if __name__ == "__main__":
    main()
