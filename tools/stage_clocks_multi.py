#!/usr/bin/env python
"""tools/stage_clocks.py for a multi-rank run (torchrun, one rank per GPU): the fused all-reduce + Adam kernel of every rank
prints when it entered, left its wait, finished its local reduction and had every rank's words (diagnostic build only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from ia2c_b200.trainer import IA2CTrainer, reference_init

def report_update_stamps(tag=""):
    """Stamps of the LAST episode's update kernels (block 0): how long each took and what the hand-overs between them cost."""
    import ctypes
    from ia2c_b200 import _lib
    lib = _lib.load()
    if not hasattr(lib, "ia2c_debug_kclocks"):
        return
    buf = (ctypes.c_ulonglong * 16)()
    lib.ia2c_debug_kclocks.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    lib.ia2c_debug_kclocks(buf)
    k = [int(v) for v in buf]
    us = lambda a, b: (k[b] - k[a]) / 1e3
    print(f"{tag}critic reduce/exchange: wait -> reduced {us(0, 1):.2f} us, -> all words in {us(0, 2):.2f}, -> block 0 done {us(0, 3):.2f}")
    print(f"{tag}hand-over to the actor gradient (block 0 done -> left the wait): {us(3, 5):.2f} us; it had entered {us(4, 5):.2f} us before")
    print(f"{tag}actor gradient: wait -> rows done {us(5, 6):.2f} us, -> block 0 done {us(5, 7):.2f}")
    print(f"{tag}hand-over to the actor reduce/exchange: {us(7, 10):.2f} us; wait -> reduced {us(10, 11):.2f}, -> all words in {us(10, 12):.2f}, -> done {us(10, 13):.2f}")
    print(f"{tag}critic exchange left its wait -> actor exchange done: {us(0, 13):.2f} us")


rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tr = IA2CTrainer(4096 * world, n_agents=2, init=reference_init(2, 5, seed=0), seed=1, rank=rank, world_size=world)
for _ in range(12):
    tr.train_episode()
torch.cuda.synchronize()
tr.check_comm()
for r in range(world):
    dist.barrier()
    if r == rank:
        report_update_stamps(f"rank {rank}: ")
        sys.stdout.flush()
dist.barrier()
dist.destroy_process_group()
