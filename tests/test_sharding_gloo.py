"""CPU tests of the N > 1 host logic with world_size 2 on gloo: contiguous sharding,
order-restoring gather of unequal blocks, max-over-ranks timing, and a sharded CPU-oracle
collision run whose gathered flags equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autonomouscar_b200 import PP_COLLISION_ORIENTED, default_params, sharding, worlds


def test_shard_bounds_cover_and_order():
    for n in (0, 1, 7, 16, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) <= max(1, sharding.padded_block(n, world))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding as orc
        p = default_params(costmap_res=0.1, costmap_width=256, costmap_height=256, collision_mode=PP_COLLISION_ORIENTED)
        grid = worlds.block_grid(256, 256, frac=0.03, block=8, seed=4)
        poses = worlds.random_poses(p, n, seed=6, inset_m=1.0)           # same on every rank
        lo, hi = sharding.shard_bounds(n, rank, world)
        local_flags, _ = orc.collide_batch(p, grid, poses[lo:hi])        # the rank's own block only
        gathered = sharding.gather_blocks(torch.from_numpy(local_flags), n)
        # 2-D records gather too (plan results are fixed-stride rows)
        rec = torch.arange(lo, hi, dtype=torch.int64).reshape(-1, 1).repeat(1, 5)
        rec_all = sharding.gather_blocks(rec, n)
        slow = sharding.max_over_ranks(1.0 + rank, torch.device("cpu"))
        # the equal-block result gather (bench.py's flag bits) on its plain-collective path: two slots, used twice
        rg = sharding.ResultGather(8, torch.int32, torch.device("cpu"), transport="symm", slots=2)
        rg_ok = rg.transport == "collective"
        for it in range(4):
            rg.send[it & 1].copy_(torch.arange(8, dtype=torch.int32) + 100 * rank + it)
            assert rg.gather(it & 1) is None
            for r in range(world):
                rg_ok = rg_ok and bool(torch.equal(rg.all[it & 1][r], torch.arange(8, dtype=torch.int32) + 100 * r + it))
        if rank == 0:
            ret["flags"] = gathered.numpy().copy()
            ret["rec_ok"] = bool(torch.equal(rec_all[:, 0], torch.arange(n)))
            ret["slow"] = slow
        ret[f"rg_ok_{rank}"] = rg_ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 4096])
def test_two_rank_gather_restores_input_order(oracle, n):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    p = default_params(costmap_res=0.1, costmap_width=256, costmap_height=256, collision_mode=PP_COLLISION_ORIENTED)
    grid = worlds.block_grid(256, 256, frac=0.03, block=8, seed=4)
    want, _ = oracle.collide_batch(p, grid, worlds.random_poses(p, n, seed=6, inset_m=1.0))
    np.testing.assert_array_equal(ret["flags"], want)
    assert 0 < want.sum() < n
    assert ret["rec_ok"] and ret["slow"] == 2.0
    assert ret["rg_ok_0"] and ret["rg_ok_1"]


def _rollout_worker(rank, world, port, n, ticks, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding as orc
        from test_rollout import _oracle_batch, _starts
        from autonomouscar_b200 import PP_COLLISION_AS_SHIPPED
        p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000)
        trace = sharding.rollout_sharded(_starts(n), _oracle_batch(orc, p, orc.MATH_DET), ticks, torch.device("cpu"))
        if rank == 0:
            ret["trace"] = trace.numpy().copy()
    finally:
        dist.destroy_process_group()


def test_two_rank_rollout_equals_single_process(oracle):
    """Cars sharded over two ranks (5 = 3 + 2), each rank driving its own block, states gathered once: the
    trace equals the single-process roll-out of all five cars."""
    from test_rollout import _oracle_batch, _starts
    from autonomouscar_b200 import PP_COLLISION_AS_SHIPPED, rollout
    n, ticks = 5, 4
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rollout_worker, args=(2, _free_port(), n, ticks, ret), nprocs=2, join=True)
    p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000)
    want = rollout.RouteRollout(_starts(n)).run(_oracle_batch(oracle, p, oracle.MATH_DET), ticks)
    np.testing.assert_array_equal(ret["trace"], want)
    assert ret["trace"].shape == (ticks, n, 4) and ret["trace"][-1, 0, 0] > ret["trace"][0, 0, 0]
