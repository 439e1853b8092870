#!/usr/bin/env python
"""BASELINE config 5: Org-N (builder-defined), 256 agents/env x 8192 envs sharded over the ranks (strong scaling),
gradient exchange by the fused NVLink all-reduce + Adam kernel.  Launch with torchrun, one rank per GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from ia2c_b200.trainer import IA2CTrainer, reference_init


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, E_total, T = int(os.environ.get("CFG5_AGENTS", 256)), int(os.environ.get("CFG5_ENVS", 8192)), 30
    steps = int(os.environ.get("CFG5_STEPS", 5))
    tr = IA2CTrainer(E_total, n_agents=N, init=reference_init(N, 5, seed=0), seed=5, rank=rank, world_size=world)
    for _ in range(2):
        tr.train_episode()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        tr.train_episode()
    b.record()
    b.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tr.check_comm()
    if rank == 0:
        per = float(ms.item()) / steps
        print(json.dumps({"config": "cfg5 Org-N (builder-defined)", "agents": N, "envs_total": E_total, "n_gpus": world, "comm": tr.comm,
                          "ms_per_episode": per, "agent_steps_per_s": E_total * N * T / (per * 1e-3), "scaling": "strong"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
