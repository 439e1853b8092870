"""Env sharding across ranks (host logic, no CUDA): SURVEY.md §8 e1.

Rank g owns the contiguous env block [g*E/G, (g+1)*E/G) with ALL agents, their beliefs and trajectory;
network parameters and Adam state are replicated.  There is no communication during the rollout.  Each
optimiser phase ends with ONE all-reduce(SUM) of a flat fp32 buffer [N, P+1] holding the locally summed
gradients — already scaled by 1/(T*E_total), the GLOBAL mean denominator — plus the loss in the last slot;
every rank then applies the identical Adam step.  Random streams are keyed by the GLOBAL env index
(env_offset + local index), so a sharded run draws exactly the numbers of the single-rank run.
"""
from __future__ import annotations


def shard_envs(num_envs_total: int, rank: int, world_size: int):
    """-> (env_offset, local_envs).  Requires an even split (the mean denominators assume it)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if num_envs_total % world_size:
        raise ValueError(f"num_envs={num_envs_total} must be divisible by world_size={world_size}")
    local = num_envs_total // world_size
    return rank * local, local


def shard_tape(tape, env_axis: int, rank: int, world_size: int):
    """Slice a replay tape (numpy / torch) along its env axis for this rank; None passes through."""
    if tape is None:
        return None
    off, n = shard_envs(tape.shape[env_axis], rank, world_size)
    index = [slice(None)] * tape.ndim
    index[env_axis] = slice(off, off + n)
    return tape[tuple(index)]


def global_agent_index(env_offset: int, local_env, agent, n_agents: int):
    """Philox counter of the action stream: (global env) * N + agent (csrc/trainer.cu, rollout_fused.cu)."""
    return (env_offset + local_env) * n_agents + agent


def belief_draw_index(env_offset: int, local_env, agent, slot, n_agents: int):
    """Philox counter of the belief stream (csrc/common.cuh: philox_belief_quad; oracle/philox.py: belief_uniforms):
    one Philox block serves FOUR modelled-other slots.  For the belief row r = (global env) * N + agent with
    K = N-1 modelled others, slots 4s .. 4s+3 share the draw at index r * ceil(K/4) + s.
    -> (draw index, word): the slot's uniform is (word + 0.5) * 2^-32."""
    kq = (n_agents + 2) // 4          # ceil((N-1)/4)
    row = (env_offset + local_env) * n_agents + agent
    return row * kq + slot // 4, slot % 4
