"""Multi-GPU parity (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_parity.py

Every rank also runs the SAME total workload on its own GPU as a single-rank trainer; the sharded run must
produce byte-identical trajectories for its env block (random streams are keyed by the global env index) and
parameters equal to the single-rank ones up to the cross-rank summation order (<= 1e-6 relative)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ia2c_b200.trainer import IA2CTrainer, reference_init  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    from ia2c_b200._lib import IA2CError
    for N, E_total, fused, comm in ((2, 64 * world, True, "p2p"), (3, 16 * world, False, "p2p"), (2, 64 * world, True, "nccl"),
                                    (5, 8 * world, True, "p2p"), (2, 64 * world, True, "p2p-multicast"), (66, 2 * world, False, "p2p"),
                                    (68, 2 * world, False, "p2p")):   # 66: per-step belief kernel, 68: whole-episode belief kernel
        init = reference_init(N, 5, seed=3)
        try:
            sharded = IA2CTrainer(E_total, n_agents=N, init=init, seed=11, rank=rank, world_size=world, fused_rollout=fused, dumps=True,
                                  comm=comm)
        except IA2CError as exc:
            if comm != "p2p-multicast":
                raise
            if rank == 0:   # the decision is collective (trainer.py: _setup_comm), so every rank skips together
                print(f"MULTICAST_UNAVAILABLE {exc}")
            continue
        assert sharded.comm == comm, (sharded.comm, comm)
        if rank == 0:
            print(f"CASE N={N} E={E_total} fused={fused} comm={sharded.comm}")
        single = IA2CTrainer(E_total, n_agents=N, init=init, seed=11, fused_rollout=fused, dumps=True)
        for ep in range(3):
            sharded.train_episode()
            single.train_episode()
        torch.cuda.synchronize()
        sharded.check_comm()
        sl = slice(sharded.env_offset, sharded.env_offset + sharded.E)
        h = lambda t: t.detach().cpu().numpy()
        for name in ("obs", "reward", "act", "partner_pred", "partner_true", "belief_dump"):
            a, b = h(getattr(sharded, name)), h(getattr(single, name))[:, sl]
            if not np.array_equal(a, b):
                ok = False
                print(f"[rank {rank}] N={N} comm={comm} {name} differs from the single-rank run")
        for name in ("actor_params", "critic_params", "actor_grad_accum"):
            a, b = h(getattr(sharded, name)).astype(np.float64), h(getattr(single, name)).astype(np.float64)
            err = np.abs(a - b).max() / np.abs(b).max()
            if err > 1e-6:
                ok = False
                print(f"[rank {rank}] N={N} comm={comm} {name} rel err {err:.2e}")
        # every rank holds identical parameters after the all-reduce + Adam
        p = sharded.actor_params.clone()
        dist.broadcast(p, 0)
        if not torch.equal(p, sharded.actor_params):
            ok = False
            print(f"[rank {rank}] N={N} parameters diverged across ranks")
        # the pipelined host-tape path (one C call per rank for p2p, torch streams for nccl): same tapes, sharded by env block
        T, K, n_ep = sharded.T, N - 1, 4
        rng = np.random.RandomState(N)
        ua = [rng.rand(T + 1, E_total, N).astype(np.float32) for _ in range(n_ep)]
        ub = [rng.rand(T + 1, E_total, N, K) for _ in range(n_ep)]
        out_s = sharded.train_episodes_host([sharded.pack_host_tape(a[:, sl], b[:, sl]) for a, b in zip(ua, ub)])
        out_1 = single.train_episodes_host([single.pack_host_tape(a, b) for a, b in zip(ua, ub)])
        torch.cuda.synchronize()
        for k in range(n_ep):
            if not np.array_equal(out_s[k]["ep_return"], out_1[k]["ep_return"][sl]):
                ok = False
                print(f"[rank {rank}] N={N} comm={comm} host pipeline: episode {k} returns differ")
        for name in ("actor_params", "critic_params"):
            a, b = h(getattr(sharded, name)).astype(np.float64), h(getattr(single, name)).astype(np.float64)
            err = np.abs(a - b).max() / np.abs(b).max()
            if err > 2e-6:
                ok = False
                print(f"[rank {rank}] N={N} comm={comm} host pipeline {name} rel err {err:.2e}")
        if sharded.episode != single.episode:
            ok = False
            print(f"[rank {rank}] episode counters differ {sharded.episode} {single.episode}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_PARITY_OK" if int(flag.item()) else "MULTIGPU_PARITY_FAILED")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
