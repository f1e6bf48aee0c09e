"""Row 8(e) on real GPUs (skipped on a one-GPU box): one process per GPU over NCCL, poses and plan queries
sharded contiguously, results gathered; the gathered flags / records equal the single-GPU run byte for byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autonomouscar_b200 import PP_COLLISION_ORIENTED, default_params, sharding, worlds

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _setup(n):
    p = default_params(costmap_res=0.1, costmap_width=512, costmap_height=512, collision_mode=PP_COLLISION_ORIENTED,
                       max_nodes=6000)
    grid = worlds.block_grid(512, 512, frac=0.02, block=8, seed=4)
    worlds.clear_disc(grid, p, 0.0, 0.0, 4.0)
    poses = worlds.random_poses(p, n, seed=6, inset_m=1.0)
    queries = worlds.forward_queries(64, seed=3, dist=(8.0, 18.0))
    return p, grid, poses, queries


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from autonomouscar_b200 import Planner
        p, grid, poses, queries = _setup(n)
        pl = Planner(p, device=rank)
        pl.set_grid(grid)
        lo, hi = sharding.shard_bounds(n, rank, world)
        local = pl.collide_dev(torch.from_numpy(poses[lo:hi]).cuda(rank))
        pl.sync()
        flags = sharding.gather_blocks(local, n)
        qlo, qhi = sharding.shard_bounds(len(queries), rank, world)
        rec = sharding.plan_records(pl.plan_batch(queries[qlo:qhi]), torch.device("cuda", rank))
        recs = sharding.gather_blocks(rec, len(queries))
        # the real thing: device-resident fixed-stride records (summary + path + profile), one all-gather over NVLink
        full = sharding.plan_batch_gather(pl, queries[qlo:qhi], len(queries), cap_path=128)
        torch.cuda.synchronize()
        # row N4 sharded: this rank's block of cars around this rank's planner, states gathered once
        from autonomouscar_b200 import rollout
        from test_rollout import _starts
        trace = sharding.rollout_sharded(_starts(37), rollout.plan_batch_cuda(pl), 5, torch.device("cuda", rank))
        if rank == 0:
            ret["flags"] = flags.cpu().numpy().copy()
            ret["recs"] = recs.cpu().numpy().copy()
            ret["full"] = full.cpu().numpy()[sharding.gathered_rows(len(queries), world)].copy()
            ret["trace"] = trace.cpu().numpy().copy()
        pl.close()
    finally:
        dist.destroy_process_group()


def test_nccl_gather_equals_single_gpu(cuda_device):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    from autonomouscar_b200 import Planner
    n = 100003
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    p, grid, poses, queries = _setup(n)
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    want = pl.collide(poses)
    np.testing.assert_array_equal(ret["flags"], want)
    want_rec = sharding.plan_records(pl.plan_batch(queries), torch.device("cpu")).numpy()
    np.testing.assert_array_equal(ret["recs"], want_rec)
    from autonomouscar_b200.capi import pp_query
    one = pl.plan_batch_dev((pp_query * len(queries))(*queries), len(queries), cap_path=128)
    pl.sync()
    np.testing.assert_array_equal(ret["full"], one.cpu().numpy())      # gathered paths byte-identical to the 1-GPU run
    assert 0 < want.sum() < n
    from autonomouscar_b200 import rollout
    from test_rollout import _starts
    np.testing.assert_array_equal(ret["trace"], rollout.RouteRollout(_starts(37)).run(rollout.plan_batch_cuda(pl), 5))
    pl.close()
