#!/usr/bin/env python
"""What does the L2 flush between timed steps leave behind?  Headline step timed by per-step CUDA events with
(a) no flush, (b) a 256 MiB write (L2 full of DIRTY lines: the step's own stores have to evict them through write-backs),
(c) the same write followed by a 256 MiB read of another buffer (L2 cold AND clean)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ia2c_b200.trainer import IA2CTrainer, reference_init

tr = IA2CTrainer(4096, n_agents=2, init=reference_init(2, 5, seed=0), seed=1)
for _ in range(50):
    tr.train_episode()
torch.cuda.synchronize()
w = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
r = torch.ones(64 << 20, dtype=torch.int32, device="cuda")


def run(mode, K=60):
    marks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for s, e in marks:
        if mode in ("write", "write+read"):
            w.zero_()
        if mode in ("write+read", "read"):
            r.sum()
        s.record()
        tr.train_episode()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) * 1e3 for s, e in marks)
    return ts[len(ts) // 2], ts[0], ts[-1]


for rep in range(2):
    for mode in ("none", "write", "write+read", "read"):
        print("%-11s median %.1f us  (min %.1f max %.1f)" % ((mode,) + run(mode)))
