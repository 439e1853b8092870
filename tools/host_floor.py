#!/usr/bin/env python
"""Host-side floor of one episode call: back-to-back episode time against the number of envs.  When the time stops
falling with E the stream is waiting for the host (python + ctypes + 4 PDL launches per episode), not for the GPU."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ia2c_b200.trainer import IA2CTrainer, reference_init

for E in (4096, 2048, 1024, 256, 32):
    tr = IA2CTrainer(E, n_agents=2, init=reference_init(2, 5, seed=0), seed=1)
    for _ in range(50):
        tr.train_episode()
    torch.cuda.synchronize()
    K = 400
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(K):
        tr.train_episode()
    e.record()
    t1 = time.perf_counter()
    e.synchronize()
    print("E=%5d  device %.1f us/episode   host enqueue %.1f us/episode" % (E, s.elapsed_time(e) / K * 1e3, (t1 - t0) / K * 1e6))
