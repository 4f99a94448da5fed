"""The path's one exchange (per-rollout observation statistics + episode counters) as ONE kernel over NVLink peer memory
(srb_rollout_stats_allreduce) against the two-step form it replaces (statistics kernel + all-reduce)."""
import os
import sys

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_single_rank_equals_the_statistics_kernel(built_lib):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 1000
    env = SRBVecEnv(helpers.fresh_clips(), n, precision=32, auto_reset=True, seed=4)
    env.reset()
    env.connect_ranks(None, 0, 1)
    act = torch.zeros(n, 22, device="cuda")
    for k in range(3):                                   # repeated calls: the scratch vector is re-zeroed by the kernel itself
        env.tracking_action(act, sigma=0.1, seed=3, step=k)
        env.step(act)
        ref = torch.zeros(304, dtype=torch.float64, device="cuda")
        env.rollout_stats(ref)
        acc = torch.full((304,), 7.0, dtype=torch.float64, device="cuda")
        env.rollout_stats_allreduce(acc)
        torch.cuda.synchronize()
        assert acc[0].item() == n
        assert torch.allclose(acc, ref, rtol=1e-12, atol=1e-9)
    assert env.xchg_status() == 0
    env.close()


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, helpers.ROOT); sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    from ipcdnnwalk_b200 import SRBVecEnv
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    n = 256 + 64 * rank                                   # ragged: the ranks hold different numbers of environments
    env = SRBVecEnv(helpers.fresh_clips(), n, precision=32, auto_reset=True, seed=10 + rank, device=rank)
    env.reset()
    env.connect_ranks(dist, rank, world)
    act = torch.zeros(n, 22, device=f"cuda:{rank}")
    ok = True
    for k in range(6):
        env.tracking_action(act, sigma=0.1, seed=3 + rank, step=k)
        env.step(act)
        ref = torch.zeros(304, dtype=torch.float64, device=f"cuda:{rank}")
        env.rollout_stats(ref)
        dist.all_reduce(ref)
        acc = torch.empty(304, dtype=torch.float64, device=f"cuda:{rank}")
        env.rollout_stats_allreduce(acc)
        torch.cuda.synchronize()
        ok = ok and bool(torch.allclose(acc, ref, rtol=1e-12, atol=1e-9)) and acc[0].item() == sum(256 + 64 * r for r in range(world))
        gathered = [torch.empty_like(acc) for _ in range(world)]
        dist.all_gather(gathered, acc)
        ok = ok and all(torch.equal(g, gathered[0]) for g in gathered)        # bit-identical on every rank
    ok = ok and env.xchg_status() == 0
    env.close()
    dist.destroy_process_group()
    out[rank] = ok


def test_two_ranks_over_peer_memory(built_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one node (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, 29533, out), nprocs=world, join=True)
        assert all(out.get(r) for r in range(world)), dict(out)
