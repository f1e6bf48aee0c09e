"""Fixed-stride plan records (pp_plan_batch_dev / pp_plan_records_unpack / sharding.plan_batch_gather).

CPU part: the record layout of include/prius_planner.h as numpy sees it, the host-side unpack helper of the C ABI
(no device needed), and the N > 1 gather on gloo (world size 2) with the CPU oracle standing in for the device
planner.  GPU part (-m gpu): the records the CUDA planner leaves on the device equal, byte for byte, what
pp_plan_batch copies to the host and what the oracle computes."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_COLLISION_ORIENTED, PP_PLAN_CAP, PP_PLAN_FOUND, PlanBatch, PlanRecords,
                                default_params, make_query, sharding, worlds)
from autonomouscar_b200.capi import pp_query
from autonomouscar_b200.results import RECORD_HEADER, record_dtype, record_stride


def records_from_results(results, cap_path):
    """Pack PlanResult-like objects into record bytes (test helper: what the kernel writes)."""
    rec = np.zeros(len(results), dtype=record_dtype(cap_path))
    rec["path_node_ids"] = -1
    for k, r in enumerate(results):
        h = rec["hdr"][k]
        for f in ("status", "n_expansions", "n_nodes", "n_open", "n_closed", "n_path", "n_profile", "n_candidates",
                  "n_collisions", "tree_digest"):
            h[f] = getattr(r, f)
        h["cap_path"] = cap_path
        h["truncated"] = int(r.n_path > cap_path)
        n, m = min(r.n_path, cap_path), min(r.n_profile, cap_path)
        rec["path_xy"][k, :n] = r.path_xy[:n]
        rec["path_node_ids"][k, :n] = r.path_node_ids[:n]
        rec["yaw_rates"][k, :m], rec["durations"][k, :m], rec["speeds"][k, :m] = r.yaw_rates[:m], r.durations[:m], r.speeds[:m]
    return rec.view(np.uint8).reshape(len(results), -1)


def test_record_layout_matches_the_c_abi():
    from autonomouscar_b200 import _lib
    L = _lib.load()
    assert RECORD_HEADER.itemsize == 64
    for cap in (0, 2, 3, 64, 127, 128, 4096):
        assert L.pp_plan_record_stride(cap) == record_stride(cap) == record_dtype(cap).itemsize
        assert record_stride(cap) % 16 == 0 and record_stride(cap) >= 64 + 44 * cap


def test_records_unpack_on_the_host(oracle):
    """pp_plan_records_unpack (pure host code) scatters records into caller arrays like pp_plan_batch fills them."""
    from autonomouscar_b200 import _lib
    L = _lib.load()
    p = default_params()
    qs = worlds.forward_queries(9, seed=2) + [make_query(300.0, 0.0, start_speed=1.0)]
    p.max_nodes = 7000
    res = [oracle.plan(p, None, q) for q in qs]
    assert {r.status for r in res} == {PP_PLAN_FOUND, PP_PLAN_CAP}
    for cap in (128, 16):                     # 16: shorter than the paths -> truncated, PP_ERR_CAPACITY
        raw = records_from_results(res, cap)
        buf = PlanBatch(len(qs), cap_path=cap)
        rc = L.pp_plan_records_unpack(raw.ctypes.data_as(C.c_void_p), len(qs), cap, buf.raw)
        longest = max(r.n_path for r in res)
        assert rc == (0 if longest <= cap else 4)
        for k, r in enumerate(res):
            assert buf.raw[k].status == r.status and buf.raw[k].n_nodes == r.n_nodes and buf.raw[k].tree_digest == r.tree_digest
            n, m = min(r.n_path, cap), min(r.n_profile, cap)
            np.testing.assert_array_equal(buf.path_xy[k, :n], r.path_xy[:n])
            np.testing.assert_array_equal(buf.path_node_ids[k, :n], r.path_node_ids[:n])
            np.testing.assert_array_equal(buf.yaw_rates[k, :m], r.yaw_rates[:m])
            np.testing.assert_array_equal(buf.speeds[k, :m], r.speeds[:m])
        view = PlanRecords(raw, cap)
        assert view.summary(0) == res[0].summary()
        np.testing.assert_array_equal(view.path_xy(0), res[0].path_xy[:cap])
    assert L.pp_plan_records_unpack(None, 3, 8, None) == 1     # PP_ERR_INVALID


class _OraclePlanner:
    """Stand-in with the two members sharding.plan_batch_gather uses, backed by the CPU oracle (CPU tensors)."""

    def __init__(self, orc, params):
        self.orc, self.params, self.torch_device = orc, params, torch.device("cpu")

    def plan_batch_dev(self, q_arr, n, cap_path, out):
        res = [self.orc.plan(self.params, None, q_arr[k], math_mode=self.orc.MATH_DET) for k in range(n)]
        out[:n] = torch.from_numpy(records_from_results(res, cap_path))
        return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, cap, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding as orc
        p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000)
        qs = worlds.forward_queries(n, seed=21)
        lo, hi = sharding.shard_bounds(n, rank, world)
        buffers = {}
        for _ in range(2):                                   # buffers are reused across calls
            allrec = sharding.plan_batch_gather(_OraclePlanner(orc, p), qs[lo:hi], n, cap_path=cap, buffers=buffers)
        if rank == 1:                                        # "only the rank that wants them copies them"
            ret["all"] = allrec.numpy().copy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 8])
def test_two_rank_record_gather_restores_input_order(oracle, n):
    """world_size 2 on gloo: every rank plans its block, ONE all-gather of the fixed-stride records; the gathered
    block holds all n plans (summary + path + profile) in input order, ragged tail padded with zero records."""
    cap = 96
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, cap, ret), nprocs=2, join=True)
    p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000)
    qs = worlds.forward_queries(n, seed=21)
    want = records_from_results([oracle.plan(p, None, q, math_mode=oracle.MATH_DET) for q in qs], cap)
    got = ret["all"]
    assert got.shape == (2 * sharding.padded_block(n, 2), record_stride(cap))
    rows = sharding.gathered_rows(n, 2)
    np.testing.assert_array_equal(got[rows], want)
    pad = np.setdiff1d(np.arange(got.shape[0]), rows)
    assert not got[pad].any()
    assert (PlanRecords(got[rows], cap).field("status") == PP_PLAN_FOUND).all()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [PP_COLLISION_AS_SHIPPED, PP_COLLISION_ORIENTED])
def test_device_records_equal_host_results_and_oracle(oracle, cuda_device, mode):
    from autonomouscar_b200 import Planner
    p = default_params(collision_mode=mode, max_nodes=6000)
    grid = worlds.block_grid(300, 300, frac=0.01, block=4, seed=9)
    worlds.clear_disc(grid, p, 0.0, 0.0, 4.0)
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    qs = worlds.forward_queries(300, seed=13) + [make_query(300.0, 5.0, start_speed=1.0)]      # the last one hits the cap
    n = len(qs)
    q_arr = (pp_query * n)(*qs)
    host = pl.plan_batch(qs, want_nodes=False, cap_path=256)
    for cap in (160, 24):
        dev = pl.plan_batch_dev(q_arr, n, cap_path=cap)
        pl.sync()
        raw = dev.cpu().numpy()
        want = records_from_results(host, cap)
        np.testing.assert_array_equal(raw, want)               # byte for byte, padding included
        again = pl.plan_batch_dev(q_arr, n, cap_path=cap, out=torch.full_like(dev, 0xAB))
        pl.sync()
        assert torch.equal(again, dev)                          # every byte of a record is written
    ref = [oracle.plan(p, grid, q, math_mode=oracle.MATH_DET) for q in qs[:40] + qs[-1:]]
    sub = pl.plan_batch_dev((pp_query * 41)(*(qs[:40] + qs[-1:])), 41, 160)
    pl.sync()                                                   # (the call is asynchronous on the handle's stream)
    np.testing.assert_array_equal(records_from_results(ref, 160), sub.cpu().numpy())
    assert host[-1].status == PP_PLAN_CAP and PlanRecords(raw, 24).field("truncated").sum() > 0
    assert pl.last_phase_ms()[4] > 0.0
    pl.close()
