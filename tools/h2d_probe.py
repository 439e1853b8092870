"""Pinned host -> device copy rate at the e2e tape size (3 MB) and above, one and two copy streams."""
import time

import torch

dev = torch.device("cuda")
for mb, streams in [(3, 1), (3, 2), (3, 3), (6, 1), (24, 1), (96, 1)]:
    n = mb << 20
    host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
    devb = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(4)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize()
    reps = 200
    for _ in range(2):
        t0 = time.perf_counter()
        for i in range(reps):
            with torch.cuda.stream(ss[i % streams]):
                devb[i % 4].copy_(host[i % 4], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{mb} MB x {reps}, {streams} stream(s): {dt / reps * 1e6:.1f} us per copy, {n * reps / dt / 1e9:.1f} GB/s")
