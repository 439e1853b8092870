#!/usr/bin/env python
"""bench.py — IA2C rollout-and-update throughput on B200 (agent-steps/s including the A2C update).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host CPU
                                                             # (oracle port; the reference is Python and its
                                                             # sources cannot travel to the GPU box)

A "step" is one IA2C episode: rollout of T=30 env steps for E envs x N agents (T+1 actor/belief evaluations)
followed by the critic phase and the actor phase (ia2c.py:62-129).  Workload at any N: BASELINE.json
configs[1] per GPU — Org domain, 2 agents, 4096 env instances per GPU (weak scaling: envs sharded across
ranks, gradients all-reduced by NCCL once per optimiser phase).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "agent-steps/sec incl. A2C update (Org domain)"
UNIT = "agent-steps/s"
T_STEPS, N_MODELS = 30, 5
# dram__bytes_read.sum + dram__bytes_write.sum of rollout_fused_kernel<2,5,1> per launch, from the ncu --set full
# capture summarised in profiles/r01_ncu_full_summary.md (the 4.5 MB trajectory it writes stays in the 126 MB L2)
NCU_ROLLOUT_DRAM_BYTES = 70656


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--agents", type=int, default=2)
    ap.add_argument("--no-fused-rollout", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-kernel-rooflines", action="store_true")
    ap.add_argument("--cpu-sample-envs", type=int, default=2048)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.marks = []

    def n_samples(self):
        try:
            return sum(1 for _ in open(self.tmp.name))
        except Exception:
            return 0

    def mark(self):
        self.marks.append(self.n_samples())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        rows = []
        for line in open(self.tmp.name):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                try:
                    rows.append((float(parts[0]), float(parts[1]), parts[3:7]))
                except ValueError:
                    pass
        os.unlink(self.tmp.name)
        lo = self.marks[0] if self.marks else 0
        hi = self.marks[1] if len(self.marks) > 1 else len(rows)
        load = rows[lo:max(hi, lo + 1)] or rows
        if not load:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(r[0] for r in load)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in load for i in range(4) if r[2][i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[1] for r in load), "reasons": reasons, "samples": len(load)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(n_agents, sample_envs, min_seconds, max_steps=None, warmup=1):
    """Times the oracle port (numpy restatement of ia2c.py's episode, oracle/loops.py) on the host cores."""
    import numpy as np
    from oracle import loops as L
    from ia2c_b200.trainer import reference_init

    actor, critic, fa = reference_init(n_agents, N_MODELS, seed=0)
    st = L.IA2CState(actor=actor, critic=critic, filter_action=fa)
    rng = np.random.RandomState(0)
    E, N, T = sample_envs, n_agents, T_STEPS

    def one():
        ua = rng.rand(T + 1, E, N).astype(np.float32)
        ub = rng.rand(T + 1, E, N, N - 1)
        L.ia2c_episode(st, E, u_act=ua, u_belief=ub)

    for _ in range(warmup):
        one()
    n, t0 = 0, time.perf_counter()
    while True:
        one()
        n += 1
        dt = time.perf_counter() - t0
        if (max_steps is not None and n >= max_steps) or (max_steps is None and dt >= min_seconds):
            break
    try:
        from threadpoolctl import threadpool_info
        blas = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        blas = 1
    return dict(value=E * N * T * n / dt, seconds=dt, episodes=n, blas_threads=blas, ms_per_step=1e3 * dt / n)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_port_rate(args.agents, args.cpu_sample_envs, min_seconds=0, max_steps=max(1, args.steps), warmup=max(1, min(args.warmup, 3)))
    sample = (f"{r['episodes']} episodes of {args.cpu_sample_envs} envs x {args.agents} agents x {T_STEPS} steps "
              f"(rollout + critic + actor update), oracle/loops.py numpy port of ia2c.py:62-129")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["episodes"],
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 nets / f64 env+belief", "data": "synthetic",
        "config": {"workload": f"Org domain, {args.agents} agents, {args.envs_per_gpu} envs per GPU, T={T_STEPS}, M={N_MODELS} (BASELINE configs[1])",
                   "cpu_sample_envs": args.cpu_sample_envs},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["blas_threads"], "kind": "port", "sample": sample,
                         "host_logical_cpus": os.cpu_count()},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ kernel rooflines
def kernel_rooflines(torch, _lib, peak_gbs):
    """Standalone streaming kernels at HBM-saturating sizes (algorithmic bytes per SURVEY.md §8 d4)."""
    lib = _lib.load()
    out = []
    dev = torch.device("cuda")

    def timeit(fn, iters=20):
        for _ in range(3):
            fn()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            e.synchronize()
            tot += s.elapsed_time(e)
        return tot / iters * 1e-3

    # (1) Org env step, thread per env, N=2 per-agent uint8 actions
    E = 1 << 24
    from ia2c_b200.org_env import OrgVecEnv
    env = OrgVecEnv(E, n_agents=2)
    act = torch.randint(0, 3, (E, 2), dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr()

    def env_step():
        _lib.check(lib.ia2c_org_step_agents(_lib.ptr(env.state), _lib.ptr(env.hist), _lib.ptr(env.cls), None, _lib.ptr(act),
                                            _lib.ptr(env.obs), None, _lib.ptr(env.reward_f32), None, None, E, 2, 0, st))
    sec = timeit(env_step)
    b = (52 + 2) * E
    out.append({"kernel": "org_step_thread_kernel", "units": f"{E} env-steps (N=2)", "bytes_per_unit": 54, "achieved": b / sec / 1e9,
                "peak": peak_gbs, "unit": "GB/s", "frac": b / sec / 1e9 / peak_gbs, "bound": "hbm", "us": sec * 1e6})
    del env, act
    # (2) belief update, dense reference layout (fp64 [R,5] in/out)
    R = 1 << 23
    fa = torch.rand(5, 3, dtype=torch.float64, device=dev)
    fa /= fa.sum(1, keepdim=True)
    lik = torch.full((R, 3), 0.1, dtype=torch.float64, device=dev)
    lik[torch.arange(R, device=dev), torch.randint(0, 3, (R,), device=dev)] = 0.8
    prev = torch.full((R, 5), 0.2, dtype=torch.float64, device=dev)
    u = torch.rand(R, dtype=torch.float64, device=dev)
    ap = torch.empty(R, dtype=torch.int64, device=dev)
    bp = torch.empty(R, 5, dtype=torch.float64, device=dev)

    def dense():
        _lib.check(lib.ia2c_belief_update_dense(_lib.ptr(fa), _lib.ptr(lik), _lib.ptr(prev), _lib.ptr(u), _lib.ptr(ap), _lib.ptr(bp),
                                                None, R, 5, 3, st))
    sec = timeit(dense)
    b = 88 * R
    out.append({"kernel": "belief_dense_kernel<5,3>", "units": f"{R} belief updates", "bytes_per_unit": 88, "achieved": b / sec / 1e9,
                "peak": peak_gbs, "unit": "GB/s", "frac": b / sec / 1e9 / peak_gbs, "bound": "hbm", "us": sec * 1e6,
                "note": "actual traffic 128 B/update (likelihood row 24 B, u 8 B, ap 8 B on top of the algorithmic 88 B)"})
    del lik, prev, u, ap, bp
    # (3) belief update, packed pairwise records (Org-N: 64 agents x 63 modelled others)
    En, N = 2048, 64
    K = N - 1
    rec = torch.zeros(En, N, K, 8, dtype=torch.uint8, device=dev)
    fan = torch.rand(N, 5, 3, dtype=torch.float64, device=dev)
    fan /= fan.sum(-1, keepdim=True)
    actn = torch.randint(0, 3, (En, N), dtype=torch.uint8, device=dev)
    partner = torch.empty(En, N, dtype=torch.uint8, device=dev)

    def pairs():
        _lib.check(lib.ia2c_belief_update_pairs(_lib.ptr(rec), _lib.ptr(fan), _lib.ptr(actn), None, None, None, _lib.ptr(partner),
                                                En, N, 5, 0, 1, 0, 1, 0, st))
    _lib.check(lib.ia2c_belief_update_pairs(_lib.ptr(rec), _lib.ptr(fan), _lib.ptr(actn), None, None, None, _lib.ptr(partner),
                                            En, N, 5, 1, 1, 0, 0, 0, st))
    sec = timeit(pairs)
    pairs_n = En * N * K
    b = 16 * pairs_n
    out.append({"kernel": "belief_pairs_kernel<5>", "units": f"{pairs_n} (agent, modelled-other) updates, N=64", "bytes_per_unit": 16,
                "achieved": b / sec / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": b / sec / 1e9 / peak_gbs, "bound": "hbm",
                "us": sec * 1e6, "pairs_per_s": pairs_n / sec,
                "note": "uint8 records (16 B/update instead of 88 B) make this kernel fp64-pipe bound, not HBM bound"})
    return out


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch

    from ia2c_b200 import _lib
    from ia2c_b200.trainer import IA2CTrainer, reference_init

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    N, E_gpu, T, M = args.agents, args.envs_per_gpu, T_STEPS, N_MODELS
    E_total = E_gpu * world
    fused = (not args.no_fused_rollout) and N <= 8
    init = reference_init(N, M, seed=0)  # identical on every rank
    tr = IA2CTrainer(E_total, n_agents=N, n_models=M, steps_per_episode=T, init=init, seed=1234, device=dev, rank=rank,
                     world_size=world, fused_rollout=fused)
    peaks, peak_src = measured_peaks()
    peak_gbs = float(peaks["hbm_gbs"])
    K, W = max(1, args.steps), max(3, args.warmup)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(W):
        tr.train_episode()
    barrier()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # ---- timed region A: K steps, L2 flushed between steps (outside the per-step events)
    if sampler:
        sampler.mark()
    e0 = [ev() for _ in range(K)]
    e1 = [ev() for _ in range(K)]
    e2 = [ev() for _ in range(K)]
    launches0 = _lib.launch_count()
    barrier()
    for i in range(K):
        flush.zero_()
        e0[i].record()
        tr.rollout()
        e1[i].record()
        tr.update()
        e2[i].record()
        tr.episode += 1
    barrier()
    launches = _lib.launch_count() - launches0
    step_ms = sum(a.elapsed_time(b) for a, b in zip(e0, e2))
    rollout_ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
    total_ms = max_over_ranks(step_ms)
    # ---- timed region B: K steps back to back, one event pair (no flush) — informational
    barrier()
    s, e = ev(), ev()
    s.record()
    for _ in range(K):
        tr.train_episode()
    e.record()
    barrier()
    b2b_ms = max_over_ranks(s.elapsed_time(e))
    # ---- e2e: the public host-buffer call — pinned uniforms H2D, episode, losses + returns D2H, sync
    n_bufs = 4
    rng = np.random.RandomState(rank)
    tapes = [tr.pack_host_tape(rng.rand(T + 1, E_gpu, N).astype(np.float32), rng.rand(T + 1, E_gpu, N, N - 1))
             for _ in range(n_bufs)]
    h2d = tapes[0].numel()
    d2h = tr._result_region.numel()

    chunk = 50   # episodes per pipelined host call (each call ends with one stream sync)

    def e2e_run(n):
        done = 0
        while done < n:
            m = min(chunk, n - done)
            tr.train_episodes_host([tapes[(done + j) % n_bufs] for j in range(m)])
            done += m

    e2e_run(3)
    barrier()
    s, e = ev(), ev()
    t0 = time.perf_counter()
    s.record()
    e2e_run(K)
    e.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(s.elapsed_time(e), e2e_wall_ms))
    # the e2e pipeline runs at the rate of its one pinned H2D copy per step: report that rate on this box next to it
    probe_dst = torch.empty_like(tapes[0], device=dev)
    for _ in range(3):
        probe_dst.copy_(tapes[0], non_blocking=True)
    s, e = ev(), ev()
    s.record()
    for j in range(20):
        probe_dst.copy_(tapes[j % n_bufs], non_blocking=True)
    e.record()
    e.synchronize()
    h2d_us = s.elapsed_time(e) / 20 * 1e3
    del probe_dst
    # keep the same load running until the clock sampler (rank 0) has data; the decision is COLLECTIVE so that every rank
    # runs the same number of extra episodes (the gradient exchange needs all ranks in every episode)
    for _ in range(50):
        need = 1 if (sampler is not None and sampler.n_samples() - sampler.marks[0] < 5) else 0
        if dist is not None:
            flag = torch.tensor([need], dtype=torch.int32, device=dev)
            dist.broadcast(flag, 0)
            need = int(flag.item())
        if not need:
            break
        for _ in range(200):
            tr.train_episode()
        torch.cuda.synchronize()
    if sampler:
        sampler.mark()
    tr.inject()  # back to the device Philox streams
    clocks = sampler.stop() if sampler else None

    units = E_total * N * T  # agent-steps per step (whole job)
    value = units * K / (total_ms * 1e-3)
    rollout_us = rollout_ms / K * 1e3
    # algorithmic HBM bytes of one rollout launch (per rank): trajectory rows + final env/belief state
    traj_bytes = (T + 1) * E_gpu * (24 + 3 * N) + T * E_gpu * 4 + E_gpu * (26 + 8 * N * (N - 1))
    if fused:   # + the critic-gradient partials the fused critic stage writes (one 148-float row per block and agent)
        traj_bytes += (E_gpu * (2 if N <= 2 else 4 if N <= 4 else 8) // 32) * N * 148 * 4
    kname = "rollout_fused_kernel (rollout + critic gradient)" if fused else "env_step_kernel + actor_step_kernel + belief_pairs_kernel (x31)"
    achieved = traj_bytes / (rollout_us * 1e-6) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 nets / f64 env+belief", "data": "synthetic",
        "config": {"workload": f"Org domain, {N} agents, {E_gpu} envs per GPU ({E_total} total), T={T}, M={M} belief models, "
                               f"rollout + critic + actor update (BASELINE configs[1] per GPU)",
                   "parallelism": f"dp{world} (env sharding, NCCL all-reduce of gradients per optimiser phase)" if world > 1 else "single GPU",
                   "l2": "256 MiB buffer written between timed steps (L2 flush); per-step CUDA events summed",
                   "sampler": "device Philox4x32-10 inverse-CDF (injected host uniforms in the e2e leg)",
                   "fused_rollout": fused, "roofline_peak_source": peak_src},
        "back_to_back_ms_per_step": b2b_ms / K,
        "e2e": {"value": units * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / K, "wall_ms_per_step": e2e_wall_ms / K,
                "h2d_copy_us_alone": h2d_us, "h2d_gbs_alone": h2d / (h2d_us * 1e-6) / 1e9,
                "api": ("IA2CTrainer.train_episodes_host -> ia2c_train_episodes_host (C ABI, pinned host tapes in, losses + returns "
                        "out per episode; one H2D and one D2H copy per episode, H2D of episode k+1 overlaps episode k; one sync per "
                        "50 episodes)") if world == 1 else
                       "IA2CTrainer.train_episodes_host -> ia2c_train_episodes_host_p2p (the same pipeline on every rank, fused NVLink "
                       "all-reduce + Adam after each gradient phase; the ranks share the host's PCIe bandwidth)"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                     "traffic": NCU_ROLLOUT_DRAM_BYTES if (fused and N == 2 and E_gpu == 4096) else None, "bytes_per_launch": traj_bytes, "us_per_launch": rollout_us,
                     "share_of_step": rollout_ms / step_ms,
                     "note": "latency-bound by construction: 31 sequential steps per env and only E*N = 8192 lanes (a 4-stage "
                             "warp-specialised pipeline, every pipe < 25 % busy, DRAM ~0: profiles/r01_ncu_full_summary.md); "
                             "the HBM-bound streaming kernels are reported under 'kernels'"},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.skip_kernel_rooflines:
        del tr, flush
        torch.cuda.empty_cache()
        line["kernels"] = kernel_rooflines(torch, _lib, peak_gbs)
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        r = cpu_port_rate(N, args.cpu_sample_envs, min_seconds=12.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["blas_threads"], "kind": "port",
                                "sample": f"{r['episodes']} episodes of {args.cpu_sample_envs} envs x {N} agents x {T} steps in {r['seconds']:.1f} s "
                                          f"(oracle/loops.py numpy port of ia2c.py:62-129)",
                                "host_logical_cpus": os.cpu_count()}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
