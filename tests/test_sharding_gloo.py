"""Multi-process (world_size 2, gloo, CPU) test of the host-side sharding / gather logic that
the N > 1 GPU path uses (phantom_b200/sharding.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phantom_b200.env import BatchStep
from phantom_b200.sharding import (LEAN_PLANES, PLANES, PackedOutputs, block_layout, gather_packed,
                                    gather_step, pack_step, shard_range, unpack_step)


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 65536, 131072 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o1, c1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def fake_step(env_ids: torch.Tensor, S=2, O=3, T=None) -> BatchStep:
    """A BatchStep whose every value encodes (global env id, agent, component)."""
    E = env_ids.numel()
    lead = () if T is None else (T,)
    base = env_ids.float().reshape(*([1] * len(lead)), E, 1, 1)
    if T is not None:
        base = base + 1000.0 * torch.arange(T).float().reshape(T, 1, 1, 1)
    obs = base + torch.arange(S).float().reshape(S, 1) * 0.1 + torch.arange(O).float() * 0.01
    ids8 = (env_ids % 200).to(torch.uint8).reshape(*([1] * len(lead)), E, 1)
    ids8 = ids8.expand(*lead, E, S).contiguous()
    return BatchStep(obs.expand(*lead, E, S, O).contiguous(), ids8, obs[..., 0].contiguous(),
                     (ids8 % 3), (ids8 % 2), ((ids8 + 1) % 2),
                     ids8[..., :2].contiguous() if S >= 2 else ids8.repeat(1, 2))


def test_pack_unpack_roundtrip():
    for T in (None, 4):
        step = fake_step(torch.arange(10), T=T)
        packed = pack_step(step)
        back = unpack_step(packed)
        for a, b in zip(step, back):
            assert torch.equal(a, b) and a.dtype == b.dtype  # u8 planes stay u8: no widening
        # ONE block: 10 envs x 2 agents x (12 + 4 + 4) bytes + 20 all_done bytes, plane aligned
        assert packed.block.dtype == torch.uint8 and packed.block.is_contiguous()
        for v in back:
            assert v.untyped_storage().data_ptr() == packed.block.untyped_storage().data_ptr()


def test_block_layout_is_aligned_and_lean_planes_are_null():
    table, nbytes = block_layout((100,), 8192, 1, 3)
    assert list(table) == list(PLANES) and nbytes % 256 == 0
    for off, _, _ in table.values():
        assert off % 256 == 0
    lean = PackedOutputs((5,), 7, 1, 3, "cpu", planes=LEAN_PLANES)
    assert lean.step.obs_mask is None and lean.step.terminations is None
    assert lean.step.observations.shape == (5, 7, 1, 3) and lean.step.all_done.shape == (5, 7, 2)
    payload = 5 * 7 * (12 + 4 + 2)
    assert payload <= lean.nbytes < payload + 3 * 256


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        offset, count = shard_range(total, rank, world)
        for T in (None, 3):
            local = fake_step(torch.arange(offset, offset + count), T=T)
            whole = gather_step(local, total)
            want = fake_step(torch.arange(total), T=T)
            for a, b in zip(whole, want):
                assert torch.equal(a, b), (rank, T)
            # the zero-copy route: results written into the packed block's views, one collective
            cap = max(shard_range(total, r, world)[1] for r in range(world))
            lead = () if T is None else (T,)
            packed = PackedOutputs(lead, count, 2, 3, "cpu", capacity_envs=cap)
            for name in PLANES:
                getattr(packed.step, name).copy_(getattr(local, name))
            got = gather_packed(packed, total, async_op=True).wait()
            for r in range(world):
                o, c = shard_range(total, r, world)
                for a, b in zip(got.rank_step(r), fake_step(torch.arange(o, o + c), T=T)):
                    assert torch.equal(a, b), (rank, r, T)
            only0 = gather_packed(packed, total, dst=0)  # gather to the trainer rank only
            assert (only0 is None) == (rank != 0)
            if rank == 0:
                for a, b in zip(only0.whole(), want):
                    assert torch.equal(a, b), (rank, T, "gather dst=0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 65])
def test_gather_step_world_size_2_gloo(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, total), nprocs=2, join=True)
