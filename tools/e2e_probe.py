#!/usr/bin/env python
"""Where does the end-to-end (host tape) episode time go?  Single GPU.

Prints: the GPU's NUMA node, the pinned H2D rate before / after binding the process to the GPU-local CPUs, the
back-to-back episode time with Philox vs device-resident injected tapes (no copies), and the pipelined host call with
the C side's enqueue / total split (IA2C_TRACE_HOST)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IA2C_TRACE_HOST"] = "1"

import numpy as np
import torch

from ia2c_b200 import hostmem
from ia2c_b200.trainer import IA2CTrainer, reference_init


def ev():
    return torch.cuda.Event(enable_timing=True)


def h2d_rate(nbytes, n=20):
    src = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
    for s_ in src:
        s_.fill_(1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        dst.copy_(src[0], non_blocking=True)
    s, e = ev(), ev()
    s.record()
    for j in range(n):
        dst.copy_(src[j % 4], non_blocking=True)
    e.record()
    e.synchronize()
    us = s.elapsed_time(e) / n * 1e3
    return us, nbytes / us / 1e3


def main():
    torch.cuda.set_device(0)
    E, N, T = 4096, 2, 30
    print("gpu numa:", hostmem.gpu_numa_cpus(0)[0], "cpus allowed:", len(os.sched_getaffinity(0)), "of", os.cpu_count())
    try:
        for n in sorted(os.listdir("/sys/devices/system/node")):
            if n.startswith("node"):
                print(" ", n, open(f"/sys/devices/system/node/{n}/cpulist").read().strip())
    except Exception as exc:
        print("no sysfs numa:", exc)
    print("H2D 3 MB unbound: %.1f us  %.1f GB/s" % h2d_rate(3047424))
    print("bind:", hostmem.bind_to_gpu_numa_node(0))
    print("H2D 3 MB bound:   %.1f us  %.1f GB/s" % h2d_rate(3047424))
    tr = IA2CTrainer(E, n_agents=N, init=reference_init(N, 5, seed=0), seed=1)
    for _ in range(20):
        tr.train_episode()
    torch.cuda.synchronize()

    def b2b(K=200):
        s, e = ev(), ev()
        s.record()
        for _ in range(K):
            tr.train_episode()
        e.record()
        e.synchronize()
        return s.elapsed_time(e) / K * 1e3

    print("b2b Philox:            %.1f us/episode" % b2b())
    # H2D on a side stream while episodes run on the main stream
    side = torch.cuda.Stream()
    src = [torch.empty(3047424, dtype=torch.uint8).pin_memory() for _ in range(4)]
    dst = torch.empty(3047424, dtype=torch.uint8, device="cuda")
    for rep in range(3):
        with torch.cuda.stream(side):
            for j in range(4):
                dst.copy_(src[j], non_blocking=True)
        torch.cuda.synchronize()
        s0, e0 = ev(), ev()
        for _ in range(100):
            tr.train_episode()
        with torch.cuda.stream(side):
            s0.record(side)
            for j in range(40):
                dst.copy_(src[j % 4], non_blocking=True)
            e0.record(side)
        torch.cuda.synchronize()
        print("H2D 3 MB while kernels run: %.1f us/copy" % (s0.elapsed_time(e0) / 40 * 1e3))
        time.sleep(0.3)
        print("H2D 3 MB idle GPU:          %.1f us  %.1f GB/s" % h2d_rate(3047424))
    rng = np.random.RandomState(0)
    ua, ub = rng.rand(T + 1, E, N).astype(np.float32), rng.rand(T + 1, E, N, N - 1)
    tr.inject(u_action=ua, u_belief=ub)
    for _ in range(5):
        tr.train_episode()
    print("b2b injected (device): %.1f us/episode" % b2b())
    print("split injected:", {k: round(v * 1e3, 1) for k, v in tr.train_episode_timed().items()})
    tapes = [tr.pack_host_tape(ua, ub) for _ in range(4)]
    for n in (50, 50, 50):
        t0 = time.perf_counter()
        tr.train_episodes_host([tapes[j % 4] for j in range(n)])
        print("host call n=%d: %.1f us/episode wall" % (n, (time.perf_counter() - t0) / n * 1e6))


if __name__ == "__main__":
    main()
