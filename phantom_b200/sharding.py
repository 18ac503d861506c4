"""Multi-GPU data parallelism for the env-step path: env instances never interact
(reference: one independent env object per Ray rollout worker, phantom/utils/rllib/train.py:
185-187,209), so the env index range is split into contiguous shards, one per rank / GPU, with
NO collective inside the step.  The RNG contract is keyed by the GLOBAL env index
(`env_offset`), so results do not depend on the number of shards.

The only communication sits after the step: the gather of every shard's obs | reward | flags
back to the trainer (the reference's rollout workers hand their results to the driver process,
phantom/utils/rllib/rollout.py:289-363).  It is ONE collective per step or per T-step rollout
over ONE packed block per rank:

  * `PackedOutputs` allocates a single contiguous byte block and hands out the seven output
    planes of the C ABI as VIEWS into it (float32 planes stay float32, uint8 planes stay
    uint8, every plane 256-byte aligned).  The step / rollout kernels write their results
    straight into those views -- there is no packing pass, no `torch.cat`, no widening.
  * `gather_packed` moves the block with `all_gather_into_tensor` (every rank gets the whole
    batch) or `gather` (only the trainer rank does) -- NCCL over NVLink / NVSwitch on GPUs, on
    the launch stream, so it is ordered after the kernel that produced the block and can
    overlap the NEXT launch (`async_op=True`).  The result is a `GatheredStep` whose planes
    are again zero-copy views, indexed [rank, ..., local env, ...].

Any torch.distributed backend works, which is how the CPU tests drive it with gloo.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

PLANES = ("observations", "obs_mask", "rewards", "reward_mask", "terminations", "truncations",
          "all_done")
LEAN_PLANES = ("observations", "rewards", "all_done")
_ALIGN = 256


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


# ------------------------------------------------------------------- the packed block
def _plane_shapes(lead: Tuple[int, ...], E: int, S: int, O: int):
    import torch

    f32, u8 = torch.float32, torch.uint8
    return {
        "observations": (lead + (E, S, O), f32), "obs_mask": (lead + (E, S), u8),
        "rewards": (lead + (E, S), f32), "reward_mask": (lead + (E, S), u8),
        "terminations": (lead + (E, S), u8), "truncations": (lead + (E, S), u8),
        "all_done": (lead + (E, 2), u8),
    }


def block_layout(lead: Tuple[int, ...], E: int, S: int, O: int,
                 planes: Sequence[str] = PLANES):
    """{plane: (byte offset, shape, dtype)} and the block size.  Planes are laid out in the
    order of `PLANES`, each at a 256-byte boundary; the layout is a pure function of the
    arguments, so every rank can compute every other rank's."""
    import math

    shapes = _plane_shapes(tuple(lead), E, S, O)
    off, table = 0, {}
    for name in PLANES:
        if name not in planes:
            continue
        shape, dtype = shapes[name]
        nbytes = math.prod(shape) * (4 if name in ("observations", "rewards") else 1)
        table[name] = (off, shape, dtype)
        off = (off + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    return table, off


def _views(block, table):
    from .env import BatchStep

    out = {}
    for name in PLANES:
        if name not in table:
            out[name] = None
            continue
        off, shape, dtype = table[name]
        n = 1
        for d in shape:
            n *= d
        nbytes = n * (4 if name in ("observations", "rewards") else 1)
        out[name] = block[off:off + nbytes].view(dtype).view(shape)
    return BatchStep(out["observations"], out["obs_mask"], out["rewards"], out["reward_mask"],
                     out["terminations"], out["truncations"], out["all_done"])


class PackedOutputs:
    """The output planes of one step (lead = ()) or one T-step rollout (lead = (T,)) of a shard,
    carved out of ONE contiguous uint8 block.  `step` is the BatchStep of views the kernels write
    (planes not in `planes` are None -> passed to the C ABI as NULL = not wanted), `block` is what
    the collective moves.  `capacity_envs` >= E sizes the block for the largest shard so that all
    ranks' blocks have one size (shards differ by at most one env)."""

    def __init__(self, lead: Tuple[int, ...], E: int, S: int, O: int, device,
                 planes: Sequence[str] = PLANES, capacity_envs: Optional[int] = None):
        import torch

        self.lead, self.E, self.S, self.O = tuple(lead), int(E), int(S), int(O)
        self.planes = tuple(p for p in PLANES if p in planes)
        self.table, used = block_layout(self.lead, self.E, self.S, self.O, self.planes)
        cap = self.E if capacity_envs is None else int(capacity_envs)
        self.nbytes = max(used, block_layout(self.lead, cap, self.S, self.O, self.planes)[1])
        self.block = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.step = _views(self.block, self.table)

    @classmethod
    def for_env(cls, env, T: Optional[int] = None, planes: Sequence[str] = PLANES,
                total_envs: Optional[int] = None, world_size: Optional[int] = None):
        import torch

        lead = () if T is None else (int(T),)
        cap = None
        if total_envs is not None and world_size:
            cap = max(shard_range(total_envs, r, world_size)[1] for r in range(world_size))
        return cls(lead, env.num_envs, max(env.spec.n_strategic, 1), env.spec.obs_dim,
                   torch.device("cuda", env.device), planes, cap)

    def launch(self, env, actions, action_mask=None):
        """phx_step / phx_rollout of `env` writing straight into this block's planes."""
        from . import _lib as L

        env._ensure_handle()
        lead = self.lead
        actions, action_mask = env._prep_actions(actions, action_mask, lead)
        p = lambda t: None if t is None else t.data_ptr()
        s = self.step
        L.check(L.lib.phx_rollout(env._handle, lead[0] if lead else 1, actions.data_ptr(),
                                  p(action_mask), p(s.observations), p(s.obs_mask), p(s.rewards),
                                  p(s.reward_mask), p(s.terminations), p(s.truncations),
                                  p(s.all_done), env._stream()))
        return s


class GatheredStep:
    """Every rank's block after the collective.  `rank_step(r)` = BatchStep of zero-copy views of
    rank r's planes ([..., E_r, ...]); `whole()` = the batch in global env order as contiguous
    tensors [..., E_total, ...] (one copy per plane, for consumers that want a single array)."""

    def __init__(self, blocks, lead, counts: List[int], S: int, O: int, planes, work=None):
        self.blocks, self.lead, self.counts = blocks, tuple(lead), list(counts)
        self.S, self.O, self.planes, self.work = S, O, tuple(planes), work

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        return self

    def rank_step(self, r: int):
        table, _ = block_layout(self.lead, self.counts[r], self.S, self.O, self.planes)
        return _views(self.blocks[r], table)

    def whole(self):
        import torch

        from .env import BatchStep

        self.wait()
        steps = [self.rank_step(r) for r in range(len(self.counts))]
        env_dim = len(self.lead)
        cat = lambda xs: None if xs[0] is None else torch.cat(xs, dim=env_dim)
        return BatchStep(*[cat([getattr(s, n) for s in steps]) for n in PLANES])


def gather_packed(packed: PackedOutputs, total_envs: int, group=None, dst: Optional[int] = None,
                  out=None, async_op: bool = False) -> Optional[GatheredStep]:
    """ONE collective over the packed block: all-gather (dst None) or gather to rank `dst`.
    `out` (uint8 [world, nbytes]) can be passed to reuse the receive buffer.  Returns the
    GatheredStep (None on the non-destination ranks of a gather)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_range(total_envs, r, world)[1] for r in range(world)]
    if counts[rank] != packed.E:
        raise ValueError("packed block does not match this rank's shard size")
    cap_bytes = max(block_layout(packed.lead, c, packed.S, packed.O, packed.planes)[1]
                    for c in counts)
    if packed.nbytes != cap_bytes:
        raise ValueError("blocks must be sized for the largest shard: build PackedOutputs with "
                         "capacity_envs / PackedOutputs.for_env(total_envs=, world_size=)")
    need_out = dst is None or rank == dst
    if need_out and out is None:
        out = torch.empty((world, packed.nbytes), dtype=torch.uint8, device=packed.block.device)
    if dst is None:
        # (flat 1-D output: the concatenation form every backend accepts)
        work = dist.all_gather_into_tensor(out.view(-1), packed.block, group=group,
                                           async_op=async_op)
    else:
        lst = [out[r] for r in range(world)] if rank == dst else None
        work = dist.gather(packed.block, lst, dst=dst, group=group, async_op=async_op)
    if not need_out:
        if async_op and work is not None:
            work.wait()
        return None
    return GatheredStep(out, packed.lead, counts, packed.S, packed.O, packed.planes,
                        work if async_op else None)


# ------------------------------------------------------- BatchStep-level convenience
def pack_step(step, out: Optional[PackedOutputs] = None, capacity_envs: Optional[int] = None):
    """Copy an existing BatchStep into a packed block (one plane copy each, dtypes kept).  The
    zero-copy route is to let the kernels write into `PackedOutputs.step` in the first place."""
    obs = step.observations
    lead, (E, S, O) = tuple(obs.shape[:-3]), obs.shape[-3:]
    if out is None:
        out = PackedOutputs(lead, E, S, O, obs.device, capacity_envs=capacity_envs)
    for name in PLANES:
        getattr(out.step, name).copy_(getattr(step, name).reshape(getattr(out.step, name).shape))
    return out


def unpack_step(packed: PackedOutputs):
    """The BatchStep of views of a packed block (inverse of pack_step)."""
    return packed.step


def gather_step(step, total_envs: int, group=None):
    """All-gather a BatchStep and return the whole batch in global env order (rank 0's envs
    first).  One collective over the packed block."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    cap = max(shard_range(total_envs, r, world)[1] for r in range(world))
    packed = pack_step(step, capacity_envs=cap)
    return gather_packed(packed, total_envs, group=group).whole()
