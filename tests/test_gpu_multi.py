"""Two-GPU check of the sharded path (skipped on a single-GPU box): each rank steps its shard
with no collective; the packed NCCL all-gather reassembles the whole batch, which must equal a
single-handle run of all envs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, total, T):
    import torch
    import torch.distributed as dist

    from phantom_b200.envs.supply_chain import SupplyChainEnv
    from phantom_b200.sharding import (LEAN_PLANES, PackedOutputs, gather_packed, gather_step,
                                       make_shard, shard_range)

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        A = np.random.RandomState(0).uniform(0, 100, size=(T, total, 1, 1)).astype(np.float32)
        off, cnt = shard_range(total, rank, world)
        env = make_shard(SupplyChainEnv, total, seed=21)
        assert (env.num_envs, env.env_offset, env.device) == (cnt, off, rank)
        env.reset_batch()
        local = env.rollout_batch(A[:, off:off + cnt])
        whole = gather_step(local, total)
        # the zero-copy route: a second episode whose kernels write straight into the packed
        # block (lean planes), one ncclAllGather of the block, overlapping nothing here
        packed = PackedOutputs.for_env(env, T, LEAN_PLANES, total_envs=total, world_size=world)
        packed.launch(env, A[:, off:off + cnt])
        got = gather_packed(packed, total, async_op=True).wait().whole()
        only0 = gather_packed(packed, total, dst=0)
        if rank == 0:
            ref = SupplyChainEnv(num_envs=total, seed=21, device=0)
            ref.reset_batch()
            want = ref.rollout_batch(A)
            for a, b in zip(whole, want):
                assert torch.equal(a.cpu(), b.cpu())
            want2 = ref.rollout_batch(A)
            for name in LEAN_PLANES:
                assert torch.equal(getattr(got, name).cpu(), getattr(want2, name).cpu()), name
                assert torch.equal(getattr(only0.whole(), name).cpu(), getattr(want2, name).cpu()), name
            assert got.obs_mask is None
            ref.close()
        env.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_shards_equal_single_handle():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 4097, 20), nprocs=2, join=True)
