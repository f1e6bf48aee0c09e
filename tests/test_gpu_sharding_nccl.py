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
        # the packed flags through ResultGather, both transports, two slots each used twice: every rank checks that
        # it holds every rank's bits (equal-sized blocks: the first `per` poses of each shard)
        per = (n // world) // 32 * 32
        mine = torch.from_numpy(poses[lo:lo + per]).cuda(rank)
        torch.cuda.synchronize()
        want_bits = pl.pack_flags_dev(pl.collide_dev(mine))
        pl.sync()
        for transport in ("symm", "nccl"):
            rg = sharding.ResultGather(per // 32, torch.int32, torch.device("cuda", rank), transport=transport, slots=2)
            for it in range(4):
                slot = it & 1
                pl.collide_bits_dev(mine, out=rg.send[slot])
                ev = torch.cuda.Event()
                ev.record(pl.stream)
                rg.gather(slot, after=ev).synchronize()
                assert torch.equal(rg.all[slot][rank], want_bits), (transport, it)
                ret[f"bits_{transport}_{it}_{rank}"] = rg.all[slot].cpu().numpy().copy()
            ret[f"transport_{transport}_{rank}"] = rg.transport
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
    # ResultGather: every rank ended up with the same [world, words] block, equal to the single-GPU bits of each shard
    per = (n // world) // 32 * 32
    for transport in ("symm", "nccl"):
        for it in range(4):
            blocks = [ret[f"bits_{transport}_{it}_{r}"] for r in range(world)]
            for r in range(world):
                lo, _ = sharding.shard_bounds(n, r, world)
                bits = np.packbits(want[lo:lo + per], bitorder="little").view(np.int32)
                for b in blocks:
                    np.testing.assert_array_equal(b[r], bits)
    assert ret["transport_nccl_0"] == "nccl"
    from autonomouscar_b200 import rollout
    from test_rollout import _starts
    np.testing.assert_array_equal(ret["trace"], rollout.RouteRollout(_starts(37)).run(rollout.plan_batch_cuda(pl), 5))
    pl.close()
