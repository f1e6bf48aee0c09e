"""GPU parity for the sampling planner (rows S1-S3 fused: Philox sample -> nearest ->
steer -> along-edge oriented collision -> append), pp_rrt_plan / _batch vs the oracle."""
import numpy as np
import pytest

from autonomouscar_b200 import (PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB, PP_PLAN_CAP, PP_PLAN_FOUND,
                                default_params, make_query, worlds)

pytestmark = pytest.mark.gpu


def _same(got, ref):
    assert got.summary() == ref.summary()
    np.testing.assert_array_equal(got.node_state.view(np.uint64), ref.node_state.view(np.uint64))
    np.testing.assert_array_equal(got.node_parent, ref.node_parent)
    np.testing.assert_array_equal(got.path_node_ids, ref.path_node_ids)
    np.testing.assert_array_equal(got.path_xy, ref.path_xy)
    np.testing.assert_array_equal(got.yaw_rates, ref.yaw_rates)
    np.testing.assert_array_equal(got.speeds, ref.speeds)


def _world(size=1024, seed=2):
    p = default_params(costmap_res=0.1, costmap_width=size, costmap_height=size, collision_mode=PP_COLLISION_ORIENTED,
                       time_step_ms=1000.0, rrt_samples_per_round=64, rrt_max_samples=60000, max_nodes=40000,
                       rrt_goal_tolerance=1.5)
    g = worlds.block_grid(size, size, frac=0.01, block=16, seed=seed, noise=False)
    worlds.clear_disc(g, p, -30.0, -30.0, 5.0)
    worlds.clear_disc(g, p, 30.0, 25.0, 5.0)
    return p, g


@pytest.mark.parametrize("K", [1, 64, 300])
def test_rrt_matches_oracle(oracle, cuda_device, K):
    from autonomouscar_b200 import Planner
    p, g = _world()
    p.rrt_samples_per_round = K
    if K == 1:      # textbook RRT, one sample per round: keep it short
        p.rrt_max_samples = 1500
    q = make_query(30.0, 25.0, start_x=-30.0, start_y=-30.0, start_heading=0.7, start_speed=1.0, seed=42, query_id=3)
    pl = Planner(p, device=0)
    pl.set_grid(g)
    got = pl.rrt_plan(q, cap_nodes=65536)
    ref = oracle.rrt_plan(p, g, q, cap_nodes=65536)
    assert ref.status == (PP_PLAN_CAP if K == 1 else PP_PLAN_FOUND)
    assert ref.n_collisions > 0 and ref.n_nodes > 500
    _same(got, ref)
    pl.close()


def test_rrt_ref_aabb_mode_and_caps(oracle, cuda_device):
    from autonomouscar_b200 import Planner
    p, g = _world(seed=5)
    p.collision_mode = PP_COLLISION_REF_AABB
    p.max_nodes = 700              # node cap bites before the goal is reached
    q = make_query(30.0, 25.0, start_x=-30.0, start_y=-30.0, start_heading=0.0, start_speed=0.5, seed=7, query_id=1)
    pl = Planner(p, device=0)
    pl.set_grid(g)
    got = pl.rrt_plan(q, cap_nodes=4096)
    ref = oracle.rrt_plan(p, g, q, cap_nodes=4096)
    assert ref.status == PP_PLAN_CAP and ref.n_path == 0 and 700 <= ref.n_nodes < 700 + 64
    _same(got, ref)
    pl.close()


def test_rrt_batch_matches_oracle(oracle, cuda_device):
    from autonomouscar_b200 import Planner
    p, g = _world(size=512, seed=11)
    p.rrt_max_samples = 8000
    p.max_nodes = 8000
    worlds.clear_disc(g, p, -15.0, -15.0, 5.0)
    rng = np.random.default_rng(3)
    qs = []
    for k in range(24):
        gx, gy = rng.uniform(-20, 20, 2)
        worlds.clear_disc(g, p, gx, gy, 4.0)
        qs.append(make_query(gx, gy, start_x=-15.0, start_y=-15.0, start_heading=rng.uniform(-3, 3), start_speed=1.0,
                             seed=99, query_id=k))
    pl = Planner(p, device=0)
    pl.set_grid(g)
    batch = pl.rrt_plan_batch(qs, want_nodes=True, cap_nodes=10000)
    n_found = 0
    for q, got in zip(qs, batch):
        ref = oracle.rrt_plan(p, g, q, cap_nodes=10000)
        _same(got, ref)
        n_found += ref.status == PP_PLAN_FOUND
    assert n_found >= 2
    # same query, different stream id -> different tree
    assert batch[0].tree_digest != batch[1].tree_digest
    pl.close()
