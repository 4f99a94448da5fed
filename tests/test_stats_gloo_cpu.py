"""world_size-2 gloo test of the multi-GPU path's host logic: envs are sharded by rank, the only
exchange is the all-reduce of observation-normaliser sufficient statistics."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, helpers.ROOT)
    from ipcdnnwalk_b200.stats import ObsStats
    rng = np.random.default_rng(0)
    full = rng.standard_normal((64, 150)).astype(np.float32) * 3 + 1
    shard = torch.from_numpy(full[rank::world]).double()          # env i lives on rank i % world
    st = ObsStats(150, "cpu")
    for _ in range(2):                                            # two rollouts
        st.acc[0] += shard.shape[0]
        st.acc[1:151] += shard.sum(0)
        st.acc[151:] += (shard * shard).sum(0)
        st.merge()
    mean, var = st.mean_var()
    ref = torch.from_numpy(full).double()
    ok = torch.allclose(mean, ref.mean(0), atol=1e-12) and torch.allclose(var, ref.var(0, unbiased=False), atol=1e-10) and st.count == 128
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_obs_stats_allreduce_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
