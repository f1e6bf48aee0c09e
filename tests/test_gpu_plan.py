"""GPU parity for the reference planner's search (rows A1-A22): pp_plan / pp_plan_batch
through the C ABI vs the CPU oracle in deterministic-math mode.  Everything is compared
bit for bit: node keys, g, cost, closing order, parent ids, expansion order, path, profile."""
import numpy as np
import pytest

from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB,
                                PP_PLAN_CAP, PP_PLAN_EXHAUSTED, PP_PLAN_FOUND, default_params, make_query, worlds)

pytestmark = pytest.mark.gpu


def _assert_same(got, ref, nodes=True):
    assert got.summary() == ref.summary()
    np.testing.assert_array_equal(got.path_xy, ref.path_xy)
    np.testing.assert_array_equal(got.path_node_ids, ref.path_node_ids)
    np.testing.assert_array_equal(got.yaw_rates, ref.yaw_rates)
    np.testing.assert_array_equal(got.durations, ref.durations)
    np.testing.assert_array_equal(got.speeds, ref.speeds)
    np.testing.assert_array_equal(got.expansion_ids, ref.expansion_ids)
    if nodes:
        # bit-exact doubles: compare the raw bit patterns (NaN-safe)
        np.testing.assert_array_equal(got.node_state.view(np.uint64), ref.node_state.view(np.uint64))
        np.testing.assert_array_equal(got.node_closed_seq, ref.node_closed_seq)
        np.testing.assert_array_equal(got.node_parent, ref.node_parent)


QUERIES = [
    dict(goal_x=20.0, goal_y=0.0, start_speed=0.0),     # SURVEY section 6 probe rows
    dict(goal_x=20.0, goal_y=0.0, start_speed=1.5),
    dict(goal_x=20.0, goal_y=3.0, start_speed=1.5),
    dict(goal_x=12.0, goal_y=-4.0, start_speed=0.7),
    dict(goal_x=8.0, goal_y=1.0, start_speed=1.5, goal_speed=0.0),   # stop at the goal: v == 0 branch of LP:205
    dict(goal_x=5.0, goal_y=0.2, start_speed=0.3, max_accel=3.0),    # 9 velocities per expansion
]


@pytest.mark.parametrize("qi", range(len(QUERIES)))
def test_plan_free_space_matches_oracle(oracle, cuda_device, qi):
    from autonomouscar_b200 import Planner
    p = default_params(max_nodes=12000)
    pl = Planner(p, device=0)
    q = make_query(**QUERIES[qi])
    got = pl.plan(q)
    ref = oracle.plan(p, None, q, math_mode=oracle.MATH_DET)
    # the decelerate-to-stop query runs into the node cap in the reference algorithm too
    assert ref.status == (PP_PLAN_CAP if QUERIES[qi].get("goal_speed") == 0.0 else PP_PLAN_FOUND)
    assert ref.n_nodes > 2000
    _assert_same(got, ref)
    pl.close()


def test_probe_numbers_from_survey(cuda_device):
    """goal (20, 0), v0 = 0 -> 4828 open + 113 closed before the
    terminal node is opened, hit point (19.9798, 0.0009)."""
    from autonomouscar_b200 import Planner
    pl = Planner(default_params(), device=0)
    r = pl.plan(make_query(20.0, 0.0, start_speed=0.0))
    assert r.status == PP_PLAN_FOUND
    assert (r.n_open, r.n_closed) == (4829, 113)
    assert abs(r.path_xy[-1, 0] - 19.9798) < 1e-4 and abs(r.path_xy[-1, 1] - 0.0009) < 1e-4
    assert r.speeds[-1] == 1.2 or abs(r.speeds[-1] - 1.2) < 1e-12
    pl.close()


def test_goal_test_int_abs_reading_matches_oracle(oracle, cuda_device):
    """A18 on the CUDA path: pp_params.goal_test_int_abs = 1 is the `int abs(int)` reading of LP:276-283."""
    from autonomouscar_b200 import Planner
    p = default_params(goal_test_int_abs=1, max_nodes=12000)
    pl = Planner(p, device=0)
    for kw in QUERIES[:4] + [dict(goal_x=5e9, goal_y=0.0, start_speed=1.0)]:
        q = make_query(**kw)
        got = pl.plan(q)
        ref = oracle.plan(p, None, q, math_mode=oracle.MATH_DET)
        _assert_same(got, ref)
    assert pl.plan(make_query(**QUERIES[0])).n_expansions < 56          # the double reading needs 56
    pl.close()


def test_caps_replace_the_wall_clock_guard(oracle, cuda_device):
    from autonomouscar_b200 import Planner
    # unreachable goal speed -> never found; node cap and expansion cap both bite
    for caps in (dict(max_nodes=3000, max_expansions=100000), dict(max_nodes=100000, max_expansions=17)):
        p = default_params(**caps)
        pl = Planner(p, device=0)
        q = make_query(200.0, 50.0, start_speed=1.0)
        got = pl.plan(q, cap_nodes=8192)
        ref = oracle.plan(p, None, q, math_mode=oracle.MATH_DET, cap_nodes=8192)
        assert ref.status == PP_PLAN_CAP and ref.n_path == 0
        _assert_same(got, ref)
        pl.close()


@pytest.mark.parametrize("mode", [PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED])
def test_plan_with_obstacles_matches_oracle(oracle, cuda_device, mode):
    from autonomouscar_b200 import Planner
    p = default_params(collision_mode=mode, max_nodes=8000)
    grid = np.zeros((300, 300), dtype=np.int8)
    # a wall across the straight line with a gap to the left (y > 0 is lower first index)
    j = 150 - int(8.0 / 0.25)
    grid[150 - 6:150 + 40, j - 2:j + 2] = 100
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    for q in (make_query(16.0, 0.0, start_speed=1.0), make_query(14.0, 2.0, start_speed=1.5)):
        got = pl.plan(q)
        ref = oracle.plan(p, grid, q, math_mode=oracle.MATH_DET)
        assert ref.n_collisions > 0
        _assert_same(got, ref)
    pl.close()


def test_frontier_exhaustion(oracle, cuda_device):
    from autonomouscar_b200 import Planner
    p = default_params(collision_mode=PP_COLLISION_REF_AABB)
    grid = np.full((300, 300), 100, dtype=np.int8)   # everything collides: no child survives
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    q = make_query(10.0, 0.0, start_speed=1.0)
    got = pl.plan(q)
    ref = oracle.plan(p, grid, q, math_mode=oracle.MATH_DET)
    assert ref.status == PP_PLAN_EXHAUSTED
    _assert_same(got, ref)
    pl.close()


def test_plan_batch_matches_single_and_oracle(oracle, cuda_device):
    from autonomouscar_b200 import Planner
    p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=6000)
    grid = worlds.block_grid(300, 300, frac=0.01, block=4, seed=9)
    worlds.clear_disc(grid, p, 0.0, 0.0, 4.0)
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    qs = worlds.forward_queries(200, seed=3)
    batch = pl.plan_batch(qs, want_nodes=False)
    secs, statuses, n_nodes, digests = oracle.plan_batch_timed(p, grid, qs, math_mode=oracle.MATH_DET, n_threads=8)
    assert [r.status for r in batch] == list(statuses)
    assert [r.n_nodes for r in batch] == list(n_nodes)
    assert [r.tree_digest for r in batch] == [int(d) for d in digests]
    assert len(set(statuses)) >= 1 and (np.array(statuses) == PP_PLAN_FOUND).sum() > 50
    # full comparison on a few of them, and batch == single
    for k in (0, 7, 63, 199):
        ref = oracle.plan(p, grid, qs[k], math_mode=oracle.MATH_DET)
        single = pl.plan(qs[k])
        _assert_same(single, ref)
        np.testing.assert_array_equal(batch[k].path_xy, ref.path_xy)
        np.testing.assert_array_equal(batch[k].speeds, ref.speeds)
    pl.close()


@pytest.mark.parametrize("mode", [PP_COLLISION_AS_SHIPPED, PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED])
def test_config1_road_world_query(oracle, cuda_device, mode):
    """BASELINE configs[0]/[1]: the reference's first plan request (next waypoint >= 20 m ahead of
    the spawn pose, local_navigator.py:586-609) on the car_demo road world rasterised at 0.1 m
    (off-road = occupied), 10k-node cap; path / parents / booleans bit-exact vs the oracle."""
    import os
    from autonomouscar_b200 import Planner
    fixture = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "road_world_4096.npz")
    grid, res, spawn = worlds.road_world(fixture)
    p = default_params(costmap_res=res, costmap_width=4096, costmap_height=4096, collision_mode=mode, max_nodes=10000)
    gx, gy = worlds.route_queries(spawn)
    assert abs(gx - 26.4527) < 1e-9 and abs(gy + 1.002) < 1e-9
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    # the straight first leg, and a goal beside the road so that children run into the verge
    for q in (make_query(gx, gy, start_speed=0.0), make_query(18.0, 6.5, start_speed=1.0)):
        got = pl.plan(q)
        ref = oracle.plan(p, grid, q, math_mode=oracle.MATH_DET)
        _assert_same(got, ref)
    assert ref.n_collisions > 0 or mode == PP_COLLISION_AS_SHIPPED
    pl.close()


def test_visited_debug_bitmap(oracle, cuda_device):
    """markVisited (LP:318-326, row A12): one mark per child returned by getNeighbors."""
    from autonomouscar_b200 import Planner
    p = default_params(collision_mode=PP_COLLISION_REF_AABB, max_nodes=4000)
    grid = np.zeros((300, 300), dtype=np.int8)
    grid[150 - 6:150 + 40, 150 - 34:150 - 30] = 100
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    q = make_query(16.0, 0.0, start_speed=1.0)
    pl.plan(q)
    ref = oracle.plan(p, grid, q, math_mode=oracle.MATH_DET, want_visited=True)
    got = pl.get_visited()
    assert ref.visited.sum() > 10
    np.testing.assert_array_equal(got, ref.visited)
    pl.close()


def test_cost_ties_follow_libstdcpp_heap_order(oracle, cuda_device):
    """Mirror-symmetric query with a yaw lattice that is exactly symmetric in binary floating
    point (heading_diff = 2^-4, heading_res = 2^-7, goal straight ahead): children at +-yaw have
    bit-identical costs, so the pop order depends on libstdc++'s push_heap / pop_heap layout
    (local_planner.hpp:111).  The GPU's fast frontier must detect the tie and reproduce it."""
    from autonomouscar_b200 import Planner
    p = default_params(heading_diff=0.0625, heading_res=0.0078125, max_nodes=6000)
    pl = Planner(p, device=0)
    for q in (make_query(20.0, 0.0, start_speed=0.0), make_query(12.0, 0.0, start_speed=1.5)):
        ref = oracle.plan(p, None, q, math_mode=oracle.MATH_DET)
        st = ref.node_state
        # there really are distinct nodes with bit-identical costs
        costs = st[1:, 7]
        assert len(np.unique(costs)) < len(costs)
        got = pl.plan(q)
        _assert_same(got, ref)
    pl.close()


def test_exact_heap_path_single_and_hybrid(oracle, cuda_device, monkeypatch):
    """The libstdc++-exact frontier is normally only used when the fast frontier detects a cost
    tie; force it (PP_PLAN_EXACT_HEAP=1) and check parity for a lone query (whole heap in shared
    memory) and for a batch large enough to use the hybrid placement (top 2048 entries in shared
    memory, deeper levels in L2) with trees of > 4000 nodes."""
    from autonomouscar_b200 import Planner
    monkeypatch.setenv("PP_PLAN_EXACT_HEAP", "1")
    p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=6000)
    grid = worlds.block_grid(300, 300, frac=0.01, block=4, seed=9)
    worlds.clear_disc(grid, p, 0.0, 0.0, 4.0)
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    qs = worlds.forward_queries(400, seed=21)
    ref0 = oracle.plan(p, grid, qs[0], math_mode=oracle.MATH_DET)
    _assert_same(pl.plan(qs[0]), ref0)
    batch = pl.plan_batch(qs, want_nodes=False)
    secs, statuses, n_nodes, digests = oracle.plan_batch_timed(p, grid, qs, math_mode=oracle.MATH_DET, n_threads=8)
    assert [r.status for r in batch] == list(statuses)
    assert [r.n_nodes for r in batch] == list(n_nodes)
    assert [r.tree_digest for r in batch] == [int(d) for d in digests]
    assert max(n_nodes) > 4000
    monkeypatch.delenv("PP_PLAN_EXACT_HEAP")
    fast = pl.plan_batch(qs, want_nodes=False)
    assert [r.tree_digest for r in fast] == [int(d) for d in digests]
    for a, b in zip(fast[:20], batch[:20]):
        np.testing.assert_array_equal(a.path_xy, b.path_xy)
        np.testing.assert_array_equal(a.speeds, b.speeds)
    pl.close()
