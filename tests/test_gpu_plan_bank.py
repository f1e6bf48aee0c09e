"""Grid banks: N simulated cars, each planning on its OWN costmap, in one pp_plan_batch_banked launch
(pp_set_grid_bank_dev builds all N derived grid representations in one set of launches).  Checked against
the single-grid path (one pp_set_grid + pp_plan_batch per grid), against the oracle, and chained behind the
batched costmap builder."""
import ctypes as C

import numpy as np
import pytest

from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB, PP_GRID_CSPACE,
                                PP_GRID_OCCUPANCY_MSG, default_params, make_query, worlds)

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert a.summary() == b.summary()
    np.testing.assert_array_equal(a.path_xy, b.path_xy)
    np.testing.assert_array_equal(a.yaw_rates, b.yaw_rates)
    np.testing.assert_array_equal(a.speeds, b.speeds)


def _grids(p, n, seed):
    out = []
    for k in range(n):
        g = worlds.block_grid(p.costmap_width, p.costmap_height, frac=[0.0, 0.01, 0.03, 0.05][k % 4], block=[2, 4, 8][k % 3],
                              seed=seed + k)
        worlds.clear_disc(g, p, 0.0, 0.0, 3.5)
        out.append(g)
    return out


@pytest.mark.parametrize("mode", [PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB])
@pytest.mark.parametrize("layout", [PP_GRID_CSPACE, PP_GRID_OCCUPANCY_MSG])
def test_banked_batch_equals_per_grid_plans(oracle, cuda_device, mode, layout):
    import torch
    from autonomouscar_b200 import Planner
    p = default_params(collision_mode=mode, max_nodes=6000)
    W, H = p.costmap_width, p.costmap_height
    n_grids = 23
    grids = _grids(p, n_grids, seed=300 + 10 * mode)
    # what the caller holds: cspace-order arrays, or the /local_costmap message order (cspace reversed, LP:437-441)
    wire = [g.reshape(-1) if layout == PP_GRID_CSPACE else g.reshape(-1)[::-1] for g in grids]
    bank = torch.from_numpy(np.ascontiguousarray(np.stack(wire))).cuda()
    pl = Planner(p, device=0)
    pl.set_grid(grids[0])                                    # the handle's own grid: must survive the bank
    pl.set_grid_bank_dev(bank.data_ptr(), n_grids, layout=layout)
    rng = np.random.default_rng(7 + mode)
    queries = [make_query(float(rng.uniform(6.0, 28.0)), float(rng.uniform(-5.0, 5.0)), start_speed=float(rng.uniform(0, 1.5)),
                          grid_index=int(rng.integers(0, n_grids)), query_id=k) for k in range(64)]
    banked = pl.plan_batch(queries, banked=True)
    own = pl.plan_batch([q for q in queries if q.reserved == 0])          # non-banked path, the handle's own grid
    for r, b in zip(own, [b for q, b in zip(queries, banked) if q.reserved == 0]):
        _same(r, b)
    single = Planner(p, device=0)
    n_coll = 0
    for gi in range(n_grids):
        mine = [(k, q) for k, q in enumerate(queries) if q.reserved == gi]
        if not mine:
            continue
        single.set_grid(grids[gi])
        for (k, _), r in zip(mine, single.plan_batch([q for _, q in mine])):
            _same(banked[k], r)
            n_coll += r.n_collisions
    assert n_coll > 0                                          # the grids mattered
    for k in (0, 1, 2):                                        # and against the oracle directly
        ref = oracle.plan(p, grids[queries[k].reserved], queries[k], math_mode=oracle.MATH_DET)
        assert banked[k].summary() == ref.summary()
        np.testing.assert_array_equal(banked[k].path_xy, ref.path_xy)
    # a smaller bank afterwards re-uses the buffers
    pl.set_grid_bank_dev(bank.data_ptr(), 5, layout=layout)
    q5 = [make_query(12.0, 1.0, grid_index=k) for k in range(5)]
    again = pl.plan_batch(q5, banked=True)
    for k in range(5):
        single.set_grid(grids[k])
        _same(again[k], single.plan(q5[k]))
    pl.close()
    single.close()


def test_bank_argument_errors(cuda_device):
    import torch
    from autonomouscar_b200 import Planner
    from autonomouscar_b200.capi import PP_ERR_INVALID, PP_ERR_NO_GRID, pp_query, pp_result
    p = default_params(collision_mode=PP_COLLISION_ORIENTED)
    pl = Planner(p, device=0)
    q = (pp_query * 1)(make_query(10.0, 0.0, grid_index=0))
    r = (pp_result * 1)()
    assert pl._L.pp_plan_batch_banked(pl._h, 1, q, r) == PP_ERR_NO_GRID          # no bank yet
    t = torch.zeros(3 * 300 * 300, dtype=torch.int8, device="cuda")
    assert pl._L.pp_set_grid_bank_dev(pl._h, 3, C.c_void_p(t.data_ptr()), 299, 300, 0.25, PP_GRID_CSPACE) == PP_ERR_INVALID
    assert pl._L.pp_set_grid_bank_dev(pl._h, 0, C.c_void_p(t.data_ptr()), 300, 300, 0.25, PP_GRID_CSPACE) == PP_ERR_INVALID
    assert pl._L.pp_set_grid_bank_dev(pl._h, 3, None, 300, 300, 0.25, PP_GRID_CSPACE) == PP_ERR_INVALID
    pl.set_grid_bank_dev(t.data_ptr(), 3)
    q[0].reserved = 3
    assert pl._L.pp_plan_batch_banked(pl._h, 1, q, r) == PP_ERR_INVALID          # grid index outside the bank
    q[0].reserved = 2
    assert pl.plan_batch([q[0]], banked=True)[0].status == 0
    pl.close()


def test_costmap_batch_feeds_a_grid_bank(oracle, cuda_device):
    """pc_build_batch -> pp_set_grid_bank_dev -> pp_plan_batch_banked, nothing leaving the device in between:
    every car's plan (ORIENTED footprint test on ITS costmap) equals the plan of the one-car chain
    pc_build -> pp_set_grid_dev -> pp_plan and the oracle's plan on the oracle's costmap."""
    from autonomouscar_b200 import LocalCostmap, Planner
    from autonomouscar_b200.costmap_capi import default_costmap_params
    cp = default_costmap_params()
    cm = LocalCostmap(cp, device=0)
    p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=20000)
    pl, one = Planner(p, device=0), Planner(p, device=0)
    n = 10
    scenes = []
    for k in range(n):
        sc = worlds.corridor_scene(70 + k, half_width=2.5 + 0.4 * (k % 4), gap_at=(9.0 + k, 12.0 + k) if k % 3 == 0 else None)
        worlds.fill_costmap_tf(sc, (3.0 * k, 0.0, 0.0))
        scenes.append(sc)
    cm.build_batch(scenes, want_host=False, want_masks=False)
    assert cm.feed_planner_bank(pl) == n
    queries = [make_query(18.0, 1.5 - 0.3 * (k % 5), start_speed=0.5, grid_index=k, query_id=k) for k in range(n)]
    banked = pl.plan_batch(queries, banked=True)
    rejected = 0
    for k, sc in enumerate(scenes):
        msg = cm.build(sc)
        cm.feed_planner(one)
        _same(banked[k], one.plan(queries[k]))
        want_msg, _ = oracle.costmap_build(cp, sc.inputs, None, math_mode=oracle.MATH_DET)
        np.testing.assert_array_equal(msg, want_msg)
        ref = oracle.plan(p, oracle.cspace_from_msg(want_msg, 300, 300), queries[k], math_mode=oracle.MATH_DET)
        assert banked[k].summary() == ref.summary()
        np.testing.assert_array_equal(banked[k].path_xy, ref.path_xy)
        rejected += ref.n_collisions
    assert rejected > 0
    cm.close()
    pl.close()
    one.close()


def test_bank_fuzz_shapes_and_parameters(oracle, cuda_device):
    """Random grid shapes (non-square), resolutions, lattice parameters and collision modes: banked plans equal the
    oracle's plans on the same grids, bit for bit."""
    import os
    import torch
    from autonomouscar_b200 import Planner
    rng = np.random.default_rng(int(os.environ.get("PP_FUZZ_SEED", "4711")))
    statuses, collisions = [], 0
    for case in range(int(os.environ.get("PP_FUZZ_CASES", "8"))):
        res = float(rng.choice([0.1, 0.2, 0.25, 0.5]))
        W, H = int(rng.integers(120, 400)), int(rng.integers(120, 400))
        p = default_params(costmap_res=res, costmap_width=W, costmap_height=H, collision_mode=int(rng.integers(1, 3)),
                           time_step_ms=float(rng.choice([150.0, 200.0])), velocity_res=float(rng.choice([0.05, 0.1])),
                           heading_diff=float(rng.choice([0.065, 0.1])), search_res=float(rng.choice([0.1, 0.25])),
                           max_nodes=int(rng.integers(300, 4000)), max_expansions=int(rng.choice([25, 100000])))
        n_grids = int(rng.integers(1, 9))
        grids = []
        for k in range(n_grids):
            g = worlds.block_grid(W, H, frac=float(rng.choice([0.0, 0.02, 0.05])), block=int(rng.choice([2, 4, 8])), seed=case * 16 + k)
            worlds.clear_disc(g, p, 0.0, 0.0, 3.5)
            grids.append(g)
        layout = int(rng.choice([PP_GRID_CSPACE, PP_GRID_OCCUPANCY_MSG]))
        wire = [g.reshape(-1) if layout == PP_GRID_CSPACE else g.reshape(-1)[::-1] for g in grids]
        bank = torch.from_numpy(np.ascontiguousarray(np.stack(wire))).cuda()
        pl = Planner(p, device=0)
        pl.set_grid_bank_dev(bank.data_ptr(), n_grids, layout=layout)
        extent = min(W, H) * res / 2
        qs = [make_query(float(rng.uniform(2.0, min(25.0, extent))), float(rng.uniform(-4.0, 4.0)),
                         goal_speed=float(rng.choice([0.0, 1.5])), start_speed=float(rng.uniform(0.0, 1.5)),
                         grid_index=int(rng.integers(0, n_grids)), query_id=k) for k in range(6)]
        for q, got in zip(qs, pl.plan_batch(qs, banked=True)):
            ref = oracle.plan(p, grids[q.reserved], q, math_mode=oracle.MATH_DET)
            assert got.summary() == ref.summary(), (case, got.summary(), ref.summary())
            np.testing.assert_array_equal(got.path_xy, ref.path_xy)
            np.testing.assert_array_equal(got.yaw_rates, ref.yaw_rates)
            statuses.append(ref.status)
            collisions += ref.n_collisions
        pl.close()
    assert len(set(statuses)) >= 2 and collisions > 0
