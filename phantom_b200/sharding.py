"""Multi-GPU data parallelism for the env-step path: env instances never interact
(reference: one independent env object per Ray rollout worker, phantom/utils/rllib/train.py:
185-187,209), so the env index range is split into contiguous shards, one per rank / GPU, with
NO collective inside the step.  The RNG contract is keyed by the GLOBAL env index
(`env_offset`), so results do not depend on the number of shards.

The only communication is optional and sits after the step: gathering the packed
obs | reward | done block of every shard to the trainer (`gather_step`), one collective per
step or per T-step rollout (NCCL over NVLink on GPUs; any torch.distributed backend works,
which is how the CPU tests drive it with gloo).
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous balanced split: (env_offset, num_envs) of `rank`.  The first
    total_envs % world_size ranks get one extra env."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(total_envs, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def make_shard(env_class, total_envs: int, *args, rank: Optional[int] = None,
               world_size: Optional[int] = None, device: Optional[int] = None, **kwargs):
    """Build this rank's shard of a `total_envs`-env batch of `env_class`.  rank / world_size
    default to torch.distributed's; device defaults to LOCAL_RANK (one process per GPU)."""
    import os

    import torch.distributed as dist

    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    offset, count = shard_range(total_envs, rank, world_size)
    return env_class(*args, num_envs=count, env_offset=offset, device=device, **kwargs)


def pack_step(step, out=None):
    """Pack a BatchStep ([..,E,S,O] obs, [..,E,S] reward / masks / flags, [..,E,2] all_done)
    into ONE contiguous float32 block [.., E, S*(O+6)+2] so that a single collective moves it."""
    import torch

    obs = step.observations
    lead, (E, S, O) = obs.shape[:-3], obs.shape[-3:]
    parts = [
        obs.reshape(*lead, E, S * O),
        step.rewards.reshape(*lead, E, S),
        step.obs_mask.reshape(*lead, E, S).float(),
        step.reward_mask.reshape(*lead, E, S).float(),
        step.terminations.reshape(*lead, E, S).float(),
        step.truncations.reshape(*lead, E, S).float(),
        step.all_done.reshape(*lead, E, 2).float(),
    ]
    width = sum(p.shape[-1] for p in parts)
    if out is None:
        out = torch.empty(*lead, E, width, dtype=torch.float32, device=obs.device)
    torch.cat(parts, dim=-1, out=out)
    return out


def unpack_step(block, S: int, O: int):
    """Inverse of pack_step."""
    import torch

    from .env import BatchStep

    lead, E = block.shape[:-2], block.shape[-2]
    sizes = [S * O, S, S, S, S, S, 2]
    obs, rew, om, rm, te, tr, ad = torch.split(block, sizes, dim=-1)
    u8 = lambda x, *shape: x.to(torch.uint8).reshape(*lead, E, *shape)
    return BatchStep(obs.reshape(*lead, E, S, O), u8(om, S), rew.reshape(*lead, E, S), u8(rm, S),
                     u8(te, S), u8(tr, S), u8(ad, 2))


def gather_step(step, total_envs: int, group=None):
    """All-gather every rank's packed step block and return the BatchStep of the whole batch in
    global env order (rank 0's envs first).  One collective; shards may differ in size by one
    env (padded to the largest shard for the collective)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    S, O = step.observations.shape[-2], step.observations.shape[-1]
    block = pack_step(step)
    env_dim = block.dim() - 2
    counts = [shard_range(total_envs, r, world)[1] for r in range(world)]
    pad = max(counts)
    if block.shape[env_dim] != counts[rank]:
        raise ValueError("step does not match this rank's shard size")
    if pad != block.shape[env_dim]:
        shape = list(block.shape)
        shape[env_dim] = pad - block.shape[env_dim]
        block = torch.cat([block, block.new_zeros(shape)], dim=env_dim)
    gathered = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(gathered, block.contiguous(), group=group)
    whole = torch.cat([g.narrow(env_dim, 0, c) for g, c in zip(gathered, counts)], dim=env_dim)
    return unpack_step(whole, S, O)
