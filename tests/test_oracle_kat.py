"""CPU tests pinning the oracle: hand-derived known answers from the cited reference lines
(the reference itself ships no tests or golden vectors -- SURVEY.md section 4 / 8(c)) and
published known-answer vectors for Philox."""
import numpy as np
import pytest

from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB,
                                PP_PLAN_CAP, PP_PLAN_EXHAUSTED, PP_PLAN_FOUND, default_params, make_query, worlds)

PHILOX_KAT = [  # Random123 kat_vectors, philox4x32 10 rounds
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_philox_known_answers(oracle):
    for ctr, key, want in PHILOX_KAT:
        assert oracle.philox4x32_10(ctr, key) == want


def test_velocity_lattice(oracle):
    # LP:191-216 with launch values: range 0.45 / 0.1 -> int(4.4999...) = 4 candidates before filtering
    p = default_params()
    q = make_query(20, 0)
    v = oracle.possible_velocities(p, q, 1.5)
    assert list(v) == [1.275, 1.375, 1.4749999999999999]          # 1.575 > max_speed dropped
    assert list(oracle.possible_velocities(p, q, 0.0)) == [(0.0 - 1.5 * 0.15) + 3 * 0.1]  # negatives dropped
    v1 = oracle.possible_velocities(p, q, 1.0)
    assert len(v1) == 4 and v1[0] == 0.775 and np.all(np.diff(v1) > 0)
    # v == 0 is skipped unless the goal speed is 0 (LP:205)
    # (with the launch values 0.225 - 1.5*0.15 is 2.8e-17, not 0: the branch needs an exact zero)
    qa = make_query(20, 0, max_accel=2.0)
    q0 = make_query(20, 0, goal_speed=0.0, max_accel=2.0)
    v_exact = 2.0 * 0.15
    assert oracle.possible_velocities(p, q0, v_exact)[0] == 0.0
    assert 0.0 not in oracle.possible_velocities(p, qa, v_exact)
    assert oracle.possible_velocities(p, q, 0.225)[0] == 0.225 - 1.5 * 0.15 > 0


def test_yaw_lattice_26_or_27(oracle):
    # LP:218-232: the loop bound is the double (h+0.065 - (h-0.065)) / 0.005, so rounding decides
    p = default_params()
    y0 = oracle.possible_yaws(p, 0.0)
    assert len(y0) == 26 and y0[0] == -0.065 and y0[1] == -0.065 + 0.005
    counts = [len(oracle.possible_yaws(p, k * 0.005)) for k in range(800)]
    assert set(counts) == {26, 27} and counts.count(27) > 0
    for h in (0.3, 1.234, -2.0):
        lo, hi = h - 0.065, h + 0.065
        n = (hi - lo) / 0.005
        want = int(np.ceil(n)) if n != int(n) else int(n)
        assert len(oracle.possible_yaws(p, h)) == want


def test_footprint_lattice_dims(oracle):
    assert oracle.lattice_dims(default_params()) == (20, 8, True)              # 0.25 m
    p01 = default_params(costmap_res=0.1, costmap_width=4096, costmap_height=4096)
    assert oracle.lattice_dims(p01) == (48, 17, True)                          # 816 samples, SURVEY S3
    assert oracle.lattice_dims(default_params(costmap_res=0.01))[2] is False   # wider than the tile window


def test_grid_flip(oracle):
    # LP:439: cspace[i][j] = data[H*W - i*H - j - 1]
    W, H = 5, 3
    msg = np.arange(W * H, dtype=np.int8)
    c = oracle.cspace_from_msg(msg, W, H)
    for i in range(W):
        for j in range(H):
            assert c[i, j] == msg[H * W - i * H - j - 1]


def _aabb_hits(oracle, p, cell_i, cell_j, pose):
    g = np.zeros((p.costmap_width, p.costmap_height), dtype=np.int8)
    g[cell_i, cell_j] = 100
    return int(oracle.collide_batch(p, g, np.array([pose], dtype=np.float32))[0][0])


@pytest.mark.parametrize("res,size,rows,cols,r0,c0", [(0.25, 300, 19, 7, 140, 146), (0.1, 4096, 47, 15, 2024, 2040)])
def test_ref_aabb_box_extent(oracle, res, size, rows, cols, r0, c0):
    """calcCollisionMatrix (LP:467-496) at the origin: 19 x 7 cells @0.25 m (rows 140..158),
    47 x 15 = 705 cells @0.1 m; the car LENGTH lies along the first (lateral) index."""
    p = default_params(costmap_res=res, costmap_width=size, costmap_height=size, collision_mode=PP_COLLISION_REF_AABB)
    pose = (0.0, 0.0, 1.0)   # yaw is ignored by the reference box
    assert _aabb_hits(oracle, p, r0, c0, pose) == 1
    assert _aabb_hits(oracle, p, r0 + rows - 1, c0 + cols - 1, pose) == 1
    assert _aabb_hits(oracle, p, r0 - 1, c0, pose) == 0
    assert _aabb_hits(oracle, p, r0 + rows, c0, pose) == 0
    assert _aabb_hits(oracle, p, r0, c0 - 1, pose) == 0
    assert _aabb_hits(oracle, p, r0, c0 + cols, pose) == 0
    # only == 100 collides (LP:451)
    g = np.full((size, size), 99, dtype=np.int8)
    assert oracle.collide_batch(p, g, np.array([pose], dtype=np.float32))[0][0] == 0


def test_oriented_rotates_with_yaw(oracle):
    p = default_params(costmap_res=0.1, costmap_width=1024, costmap_height=1024, collision_mode=PP_COLLISION_ORIENTED)
    g = np.zeros((1024, 1024), dtype=np.int8)
    # an obstacle 2 m ahead (x = +2 m -> second index 512 - 20), on the centre line
    g[512, 512 - 20] = 100
    f = oracle.collide_batch(p, g, np.array([[0, 0, 0.0], [0, 0, np.pi / 2], [0, 0, np.pi]], dtype=np.float32))[0]
    assert list(f) == [1, 0, 1]   # length 4.7 m covers +-2.35 m along the heading, width only +-0.79 m
    p.collision_mode = PP_COLLISION_AS_SHIPPED
    assert oracle.collide_batch(p, g, np.array([[0, 0, 0.0]], dtype=np.float32))[0][0] == 0   # LP:445


def test_det_math_accuracy(oracle):
    rng = np.random.default_rng(0)
    x = rng.uniform(-100, 100, 400000)
    s, c = oracle.det_sincos(x)

    def ulps(a, b):
        return np.max(np.abs(a - b) / np.spacing(np.abs(b)))
    assert ulps(s, np.sin(x)) <= 2 and ulps(c, np.cos(x)) <= 2
    y, xx = rng.uniform(-30, 30, 400000), rng.uniform(-30, 30, 400000)
    assert ulps(oracle.det_atan2(y, xx), np.arctan2(y, xx)) <= 4
    sp_y = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 0.0, 2.0, -2.0])
    sp_x = np.array([0.0, 0.0, -0.0, -0.0, 0.0, 0.0, -1.0, 2.0, -2.0])
    np.testing.assert_allclose(oracle.det_atan2(sp_y, sp_x), np.arctan2(sp_y, sp_x), rtol=0, atol=1e-15)
    d = rng.uniform(-50, 50, 1000)
    w = oracle.det_wrap_pi(d)
    assert np.all(np.abs(w) <= np.pi + 1e-12)
    np.testing.assert_allclose(np.sin(w), np.sin(d), atol=1e-12)


def test_planner_probe_rows(oracle):
    """SURVEY.md section 6 probe, re-derived by the real oracle: goal (20,0) v0=0 -> 4828 open
    + 113 closed (+1: the terminal node opened at LP:45), hit point (19.9798, 0.0009), final
    profile speed 1.2; (20,0) and (20,3) at v0=1.5 -> 3782 + 93."""
    p = default_params()
    for mm in (oracle.MATH_LIBM, oracle.MATH_DET):
        r = oracle.plan(p, None, make_query(20, 0, start_speed=0.0), math_mode=mm)
        assert r.status == PP_PLAN_FOUND and (r.n_open, r.n_closed) == (4829, 113)
        assert abs(r.path_xy[-1, 0] - 19.9798) < 1e-4 and abs(r.path_xy[-1, 1] - 0.0009) < 1e-4
        assert abs(r.speeds[-1] - 1.2) < 1e-12 and np.all(r.durations == 150.0)
        assert r.path_xy[0].tolist() == [0.0, 0.0]
        for goal in ((20, 0), (20, 3)):
            r = oracle.plan(p, None, make_query(*goal, start_speed=1.5), math_mode=mm)
            assert r.status == PP_PLAN_FOUND and (r.n_open, r.n_closed) == (3783, 93)


def test_planner_structure_is_math_library_independent(oracle):
    """glibc and the deterministic sin/cos/atan2 differ by <= a few ulp; the search tree
    (closing order, parents, expansion order, path node ids) must not depend on that for
    non-degenerate queries."""
    p = default_params()
    for q in worlds.forward_queries(12, seed=5):
        a = oracle.plan(p, None, q, math_mode=oracle.MATH_LIBM)
        b = oracle.plan(p, None, q, math_mode=oracle.MATH_DET)
        assert a.status == b.status and a.n_nodes == b.n_nodes and a.n_expansions == b.n_expansions
        np.testing.assert_array_equal(a.expansion_ids, b.expansion_ids)
        np.testing.assert_array_equal(a.node_parent, b.node_parent)
        np.testing.assert_array_equal(a.node_closed_seq, b.node_closed_seq)
        np.testing.assert_array_equal(a.path_node_ids, b.path_node_ids)
        np.testing.assert_allclose(a.node_state, b.node_state, rtol=1e-12, atol=1e-12)


def test_planner_invariants(oracle):
    p = default_params()
    r = oracle.plan(p, None, make_query(15, -2, start_speed=1.0))
    st = r.node_state
    assert r.status == PP_PLAN_FOUND
    # two closeNode() per iteration (LP:41 and LP:80): the last iteration returns before the second
    assert r.n_closed == 2 * r.n_expansions - 1
    assert r.n_open + r.n_closed == r.n_nodes
    # start node: child == parent == origin, cost 9999999999 (LP:90-95); parent id of its children is 0
    assert st[0, :4].tolist() == [0, 0, 0, 0] and st[0, 7] == 9999999999.0
    assert r.node_parent[1] == 0 and r.expansion_ids[0] == 0
    # g is the edge length only (LP:162), never cumulative
    np.testing.assert_allclose(st[1:, 6], np.hypot(st[1:, 0] - st[1:, 2], st[1:, 1] - st[1:, 3]), rtol=1e-15)
    # un-halved "average" velocity (LP:155): step = (v_parent + v_child) * 0.15
    par = r.node_parent[1:-1]
    np.testing.assert_allclose(st[1:-1, 6], (st[par, 5] + st[1:-1, 5]) * 0.15, rtol=1e-9)
    # path ends at the terminal node, profile has n_path - 2 entries (LP:548)
    assert r.n_profile == r.n_path - 2
    assert r.path_node_ids[-1] == r.n_nodes - 1


def test_planner_caps_and_exhaustion(oracle):
    p = default_params(max_expansions=9)
    r = oracle.plan(p, None, make_query(100, 0, start_speed=1.0))
    assert r.status == PP_PLAN_CAP and r.n_expansions == 9 and r.n_path == 0
    p = default_params(collision_mode=PP_COLLISION_REF_AABB)
    r = oracle.plan(p, np.full((300, 300), 100, np.int8), make_query(10, 0, start_speed=1.0))
    assert r.status == PP_PLAN_EXHAUSTED and r.n_nodes == 1 and r.n_collisions == r.n_candidates > 0


def test_goal_test_int_abs_reading(oracle):
    """A18: with `int abs(int)` (the overload the reference's mandated GCC 5 toolchain may pick at LP:276-283)
    every difference is truncated toward zero first, so the 0.1 m / 0.5 m/s tolerances act as "< 1.0": the goal
    test fires on the first extrapolated point within 1 m of the goal, at any speed within 1 m/s of the goal's.
    The double reading is the default (what g++ 13 and oracle/_ref resolve to)."""
    q = make_query(20, 0, start_speed=0.0)
    dbl = oracle.plan(default_params(), None, q)
    iab = oracle.plan(default_params(goal_test_int_abs=1), None, q)
    assert dbl.status == iab.status == PP_PLAN_FOUND
    assert iab.n_expansions < dbl.n_expansions and iab.n_nodes < dbl.n_nodes
    end_d, end_i = dbl.path_xy[-1], iab.path_xy[-1]
    assert abs(end_d[0] - 20) <= 0.1 and abs(end_d[1]) <= 0.1
    assert 0.1 < abs(end_i[0] - 20) < 1.0 and abs(end_i[1]) < 1.0          # inside the truncated box, outside the real one
    # the searches are identical up to the expansion at which the int reading stops
    np.testing.assert_array_equal(iab.expansion_ids, dbl.expansion_ids[: iab.n_expansions])
    n = iab.n_nodes - 1                                                      # (the terminal node differs)
    np.testing.assert_array_equal(iab.node_state[:n, :6], dbl.node_state[:n, :6])
    # out-of-range differences never pass (defined behaviour where the conversion would be UB)
    far = oracle.plan(default_params(goal_test_int_abs=1, max_expansions=3), None, make_query(5e9, 0, start_speed=1.0))
    assert far.status == PP_PLAN_CAP


def test_sampling_planner_oracle(oracle):
    p = default_params(costmap_res=0.1, costmap_width=1024, costmap_height=1024, collision_mode=PP_COLLISION_ORIENTED,
                       time_step_ms=1000.0, rrt_samples_per_round=64, rrt_max_samples=200000, max_nodes=100000,
                       rrt_goal_tolerance=1.5)
    g = worlds.block_grid(1024, 1024, frac=0.01, block=16, seed=2, noise=False)
    worlds.clear_disc(g, p, -30.0, -30.0, 5.0)
    worlds.clear_disc(g, p, 30.0, 25.0, 5.0)
    q = make_query(30.0, 25.0, start_x=-30.0, start_y=-30.0, start_heading=0.7, start_speed=1.0, seed=42, query_id=3)
    r = oracle.rrt_plan(p, g, q, cap_nodes=120000)
    assert r.status == PP_PLAN_FOUND and r.n_path > 5
    ids = r.path_node_ids
    assert ids[0] == 0 and np.all(r.node_parent[ids[1:]] == ids[:-1])      # the path is a parent chain
    assert abs(r.path_xy[-1, 0] - 30.0) <= 1.5 and abs(r.path_xy[-1, 1] - 25.0) <= 1.5
    # every edge obeys the motion model: heading change within +-heading_diff, step = (v + v') * dt
    st = r.node_state[: r.n_nodes]
    par = r.node_parent[1: r.n_nodes]
    dh = st[1:, 4] - st[par, 4]
    assert np.all(np.abs(dh) <= p.heading_diff + 1e-12)
    np.testing.assert_allclose(st[1:, 6], (st[par, 5] + st[1:, 5]) * 1.0, rtol=1e-12)
    # same seed -> same tree; different query_id -> different stream
    r2 = oracle.rrt_plan(p, g, q, cap_nodes=120000)
    assert r2.tree_digest == r.tree_digest
    q.query_id = 4
    assert oracle.rrt_plan(p, g, q, cap_nodes=120000).tree_digest != r.tree_digest


def test_bucket_grid_nearest_equals_the_scan(oracle):
    """S2 on large trees: the oracle's bucket-grid query (BucketGridNN) returns exactly what the linear scan returns --
    minimum float32 distance, ties -> lowest index -- on uniform sets, a blob in a corner with far-away queries (the
    early phase of a config-4 run), duplicates, points on cell borders and points outside the grid."""
    rng = np.random.default_rng(7)
    box = (-40.0, 60.0, -30.0, 50.0)
    cases = []
    cases.append((rng.uniform((-40, -30), (60, 50), (5000, 2)), rng.uniform((-40, -30), (60, 50), (3000, 2))))
    blob = rng.normal((-35.0, -25.0), 1.5, (3000, 2))
    cases.append((blob, rng.uniform((-40, -30), (60, 50), (3000, 2))))                       # far queries, tree in a corner
    lattice = np.stack(np.meshgrid(np.arange(-40, 60, 100 / 64), np.arange(-30, 50, 80 / 64)), -1).reshape(-1, 2)
    cases.append((np.concatenate([lattice, lattice]), np.concatenate([lattice[::7] + 0.5 * 100 / 64, lattice[::5]])))  # ties + borders
    cases.append((rng.uniform(-200, 200, (4000, 2)), rng.uniform(-200, 200, (2000, 2))))     # mostly outside the grid
    cases.append((np.array([[1.0, 2.0]]), rng.uniform(-50, 70, (100, 2))))                    # a single node
    for nodes, queries in cases:
        for cells in (1, 7, 64, 200):
            a_idx, a_d = oracle.nearest(nodes, queries)
            b_idx, b_d = oracle.nearest_grid(nodes, queries, box, cells)
            np.testing.assert_array_equal(a_idx, b_idx)
            np.testing.assert_array_equal(a_d, b_d)
