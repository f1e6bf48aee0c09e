"""Row N2 on the GPU: pc_build (csrc/pc_costmap.cu) through the C ABI against the oracle
(deterministic math, byte for byte) and against the committed outputs of the reference binary."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_costmap_golden as mcg  # noqa: E402


def _fill_tf(sc, car):
    from autonomouscar_b200 import worlds
    worlds.fill_costmap_tf(sc, car)


def test_costmap_bit_exact_vs_oracle(oracle, cuda_device):
    from autonomouscar_b200 import LocalCostmap, worlds
    p = mcg.params()
    cm = LocalCostmap(p, device=0)
    # PP_FUZZ_SEED / PP_FUZZ_CASES widen the sweep for one-off soak runs
    seed = int(os.environ.get("PP_FUZZ_SEED", "3"))
    rng = np.random.default_rng(seed)
    for trial in range(int(os.environ.get("PP_FUZZ_CASES", "6"))):
        car = (float(rng.uniform(-150, 150)), float(rng.uniform(-150, 150)), float(rng.uniform(-3.1, 3.1)))
        sc = worlds.costmap_scene(100 * seed + trial, n_planar=int(rng.choice([640, 333])), per_ring=int(rng.choice([512, 100])))
        _fill_tf(sc, car)
        g = mcg.global_grid(p, car, int(rng.integers(0, 160))) if trial % 3 != 1 else None
        cm.set_global_grid(g)
        got = cm.build(sc)
        want, _ = oracle.costmap_build(p, sc.inputs, g, math_mode=oracle.MATH_DET)
        np.testing.assert_array_equal(got, want)
        assert (got == 100).sum() > 200
        again = cm.build(sc)                      # graph re-launch on the same handle
        np.testing.assert_array_equal(again, want)
    assert cm.launch_count() > 0
    cm.close()


def test_costmap_matches_reference_fixture(cuda_device):
    """Against what the REFERENCE binary published (glibc math): the only differences allowed are
    cells where a last-ulp difference of sin/cos/atan2 moves a point across a cell border."""
    from autonomouscar_b200 import LocalCostmap
    z = np.load(os.path.join(GOLD, "ref_costmap.npz"))
    p = mcg.params()
    cm = LocalCostmap(p, device=0)
    for k, (seed, car, offset) in enumerate(mcg.CASES):
        sc = mcg.scene_from_fixture(z, k)
        cm.set_global_grid(mcg.global_grid(p, car, offset))
        got = cm.build(sc)
        want = z[f"c{k}_grid"]
        assert (got != want).mean() < 1e-3, (k, int((got != want).sum()))
        assert abs(int((got == 100).sum()) - int((want == 100).sum())) <= 8
    cm.close()


def test_costmap_edge_cases(oracle, cuda_device):
    from autonomouscar_b200 import LocalCostmap, worlds
    from autonomouscar_b200.costmap_capi import CostmapScene
    p = mcg.params()
    cm = LocalCostmap(p, device=0)
    # nothing received yet: every cell unknown (clearMap only)
    empty = CostmapScene(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros((0, 3), np.float32),
                         (np.float32(0), np.float32(0), np.float32(30)))
    _fill_tf(empty, (0.0, 0.0, 0.0))
    got = cm.build(empty)
    assert (got == -1).all()
    # points far outside the window are ignored (the reference would index out of bounds), NaN / inf ranges too
    sc = worlds.costmap_scene(9)
    _fill_tf(sc, (10.0, 5.0, 1.0))
    sc.right[:50] = 500.0
    sc.right[50:60] = np.nan
    sc.left[:10] = np.inf
    sc.cloud[:100, 0] = 1e6
    sc.cloud[100:110, 1] = np.nan
    sc.inputs.right.range_max = np.float32(1000.0)
    got = cm.build(sc)
    want, _ = oracle.costmap_build(p, sc.inputs, None, math_mode=oracle.MATH_DET)
    np.testing.assert_array_equal(got, want)
    cm.close()


def test_costmap_feeds_planner_on_device(oracle, cuda_device):
    """/local_costmap never leaves the GPU: pc_grid_dev -> pp_set_grid_dev, then the planner's
    m_c_space equals the LP:439 flip of the grid the costmap node would have published."""
    from autonomouscar_b200 import LocalCostmap, Planner, default_params, worlds
    from autonomouscar_b200.costmap_capi import default_costmap_params
    cm = LocalCostmap(default_costmap_params(), device=0)
    sc = worlds.costmap_scene(5)
    _fill_tf(sc, (0.0, 0.0, 0.0))
    msg = cm.build(sc)
    pl = Planner(default_params(), device=0)
    cm.feed_planner(pl)
    np.testing.assert_array_equal(pl.get_cspace(), oracle.cspace_from_msg(msg, 300, 300))
    pl.close()
    cm.close()


# ---- row N3: the global road map ---------------------------------------------------------------
import make_global_costmap_golden as mgg  # noqa: E402


def test_global_road_map_bit_exact_small(oracle, cuda_device):
    """pg_build against the oracle (deterministic math) and against the grids the REFERENCE published."""
    from autonomouscar_b200 import GlobalRoadMap
    z = np.load(os.path.join(GOLD, "ref_global_costmap.npz"))
    for k, (seed, lanes, width) in enumerate(mgg.SMALL_CASES):
        p = mgg.small_params(lanes, width)
        gm = GlobalRoadMap(z[f"s{k}_xy"], z[f"s{k}_off"], params=p, device=0)
        want, _ = oracle.global_costmap_build(p, z[f"s{k}_xy"], z[f"s{k}_off"], math_mode=oracle.MATH_DET)
        np.testing.assert_array_equal(gm.host, want)
        assert (gm.host != z[f"s{k}_grid"]).mean() < 1e-4        # the reference binary itself (glibc math)
        assert gm.launches >= 3
        gm.close()
    # degenerate inputs: no polylines / single-point polylines -> all 99 except the never-written head
    p = mgg.small_params(2, 11.1)
    gm = GlobalRoadMap(np.zeros((3, 2), np.float32), np.array([0, 1, 2, 3], np.int32), params=p, device=0)
    assert (gm.host[: mgg.SMALL + 1] == -1).all() and (gm.host[mgg.SMALL + 1:] == 99).all()
    gm.close()


def test_global_road_map_launch_file_size(cuda_device):
    """12000 x 12000 from the reference's own road polylines: the grid the reference published
    (SHA-256 / per-row / per-column road-cell counts from oracle/_ref), then handed to the local
    costmap builder without leaving the device."""
    from autonomouscar_b200 import GlobalRoadMap, LocalCostmap, worlds
    from autonomouscar_b200.costmap_capi import default_costmap_params
    z = np.load(os.path.join(GOLD, "ref_global_costmap.npz"))
    xy, off = mgg.road_edges()
    gm = GlobalRoadMap(xy, off, device=0)
    sha, rows, cols, total = mgg.full_summary(gm.host, 12000)
    assert np.abs(rows - z["full_rows"]).sum() <= 64 and np.abs(cols - z["full_cols"]).sum() <= 64
    assert abs(int(total[0]) - int(z["full_sum"][0])) <= 99 * 64
    exact = bool(np.array_equal(sha, z["full_sha"]))
    print("global map: %.2f ms on the device, identical to the reference's bytes: %s" % (gm.build_ms, exact))
    host = gm.host
    cm = LocalCostmap(default_costmap_params(), device=0)
    cm.adopt_global_grid(gm)
    car = (30.0, 10.0, 0.3)                      # on the reference's route, inside the road network
    sc = worlds.costmap_scene(2)
    _fill_tf(sc, car)
    got = cm.build(sc)
    cm2 = LocalCostmap(default_costmap_params(), device=0)
    cm2.set_global_grid(host)
    np.testing.assert_array_equal(got, cm2.build(sc))
    assert ((got > 0) & (got < 99)).sum() > 1000  # road costs made it into the local grid
    cm.close()
    cm2.close()


def test_stack_rollout_on_device(oracle, cuda_device):
    """Rows N1-N3 chained with the hot path: global road map -> local costmap -> planner, the grid
    never leaving the GPU, along the reference's own route; every cycle's plan equals the oracle's
    plan for the same goal (as shipped: the grid does not influence the search, LP:445)."""
    from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_PLAN_FOUND, GlobalRoadMap, LocalCostmap, Planner,
                                    default_params, make_query, worlds)
    from autonomouscar_b200.costmap_capi import default_costmap_params
    xy, off = mgg.road_edges()
    gm = GlobalRoadMap(xy, off, device=0, want_host=False)
    cm = LocalCostmap(default_costmap_params(), device=0)
    cm.adopt_global_grid(gm)
    p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000)
    pl = Planner(p, device=0)
    route = worlds.densified_route()
    x, y, yaw, v, wp = 0.0, 15.0, 0.0, 0.0, 0
    road_cells = 0
    for k in range(12):
        sc = worlds.costmap_scene(70 + k)
        worlds.fill_costmap_tf(sc, (x, y, yaw))
        msg = cm.build(sc, want_host=True)
        road_cells += int(((msg > 0) & (msg < 99)).sum())
        cm.feed_planner(pl)
        np.testing.assert_array_equal(pl.get_cspace(), oracle.cspace_from_msg(msg, 300, 300))
        wp, (wx, wy) = worlds.next_waypoint(route, wp, x, y)
        c, s = np.cos(yaw), np.sin(yaw)
        q = make_query(c * (wx - x) + s * (wy - y), -s * (wx - x) + c * (wy - y), start_speed=v)
        r = pl.plan(q)
        ref = oracle.plan(p, None, q, math_mode=oracle.MATH_DET)
        assert r.status == PP_PLAN_FOUND and r.summary() == ref.summary()
        np.testing.assert_array_equal(r.path_xy, ref.path_xy)
        np.testing.assert_array_equal(r.yaw_rates, ref.yaw_rates)
        v = float(r.speeds[0])
        yaw += float(r.yaw_rates[0]) * 0.05
        x += v * 0.05 * np.cos(yaw)
        y += v * 0.05 * np.sin(yaw)
    assert x > 0.02, x                      # the car moved (it starts from rest, 12 cycles of 50 ms)
    assert road_cells > 12 * 1000, road_cells   # the road overlay reached every local grid
    cm.close()
    pl.close()


def test_stack_rollout_oriented_corridor(oracle, cuda_device):
    """The same chain with the collision check switched on (ORIENTED footprint): the lasers see two walls,
    pc_build inflates them, the grid goes to the planner on the device, and the plan that avoids them equals
    the oracle's plan on the oracle's costmap of the same messages -- every cycle, bit for bit."""
    from autonomouscar_b200 import (PP_COLLISION_ORIENTED, PP_PLAN_FOUND, LocalCostmap, Planner, default_params,
                                    make_query, worlds)
    from autonomouscar_b200.costmap_capi import default_costmap_params
    cp = default_costmap_params()
    cm = LocalCostmap(cp, device=0)
    p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=20000)
    pl = Planner(p, device=0)
    route = worlds.densified_route()
    x, y, yaw, v, wp = 0.0, 12.6, 0.0, 0.5, 0          # 1.4 m to the right of the route: goals pull towards a wall
    collisions = found = 0
    for k in range(10):
        sc = worlds.corridor_scene(90 + k, half_width=2.5)
        worlds.fill_costmap_tf(sc, (x, y, yaw))
        msg = cm.build(sc, want_host=True)
        want_msg, _ = oracle.costmap_build(cp, sc.inputs, None, math_mode=oracle.MATH_DET)
        np.testing.assert_array_equal(msg, want_msg)
        assert (msg == 100).sum() > 300
        cm.feed_planner(pl)
        wp, (wx, wy) = worlds.next_waypoint(route, wp, x, y)
        c, s = np.cos(yaw), np.sin(yaw)
        q = make_query(c * (wx - x) + s * (wy - y), -s * (wx - x) + c * (wy - y), start_speed=v)
        r = pl.plan(q)
        ref = oracle.plan(p, oracle.cspace_from_msg(want_msg, 300, 300), q, math_mode=oracle.MATH_DET)
        assert r.summary() == ref.summary()
        np.testing.assert_array_equal(r.path_xy, ref.path_xy)
        np.testing.assert_array_equal(r.yaw_rates, ref.yaw_rates)
        collisions += r.n_collisions
        if r.status == PP_PLAN_FOUND and r.n_profile > 0:
            found += 1
            v = float(r.speeds[0])
            yaw += float(r.yaw_rates[0]) * 0.05
            x += v * 0.05 * np.cos(yaw)
            y += v * 0.05 * np.sin(yaw)
    assert found >= 8 and collisions > 0, (found, collisions)
    print("corridor roll-out: %d plans found, %d candidate poses rejected by the footprint test" % (found, collisions))
    cm.close()
    pl.close()


def test_gpu_chain_equals_reference_node_chain(cuda_device):
    """Two REFERENCE nodes in a row (local_costmap.cpp -> local_planner.cpp with its collision code enabled,
    both compiled from the reference sources: oracle/_ref) against the two CUDA components in a row
    (pc_build -> pp_plan, REF_AABB), fed the same sensor messages and LocalNav goals for several 20 Hz cycles:
    same published grid, same plan outcome and tree size, /local_path within 1e-9 m, /mp speeds identical."""
    from autonomouscar_b200 import (PP_COLLISION_REF_AABB, PP_PLAN_FOUND, LocalCostmap, Planner, default_params,
                                    make_query, worlds)
    from oracle import ref_binding as ref
    if not (ref.available() and ref.costmap_available()):
        pytest.skip("oracle/_ref not present")
    cp = mcg.params()
    p = default_params(collision_mode=PP_COLLISION_REF_AABB, max_nodes=20000, max_expansions=2000)
    route = worlds.densified_route()
    x, y, yaw, v, wp = 0.0, 13.2, 0.0, 0.5, 0
    ref.costmap_set_tfs(worlds.costmap_tfs(x, y, yaw))
    ref_costmap, ref_planner = ref.RefCostmap(cp), ref.RefNode(p, ref.AABB)
    cm, pl = LocalCostmap(cp, device=0), Planner(p, device=0)
    rejected = found = 0
    for k in range(6):
        sc = worlds.corridor_scene(90 + k, half_width=3.5)
        ref.costmap_set_tfs(worlds.costmap_tfs(x, y, yaw))
        ref.costmap_derive_inputs(sc)                       # the tf numbers exactly as the reference node reads them
        ref_grid, _ = ref_costmap.build(sc)
        grid = cm.build(sc, want_host=True)
        assert (grid != ref_grid).sum() <= 2, int((grid != ref_grid).sum())
        ref_planner.set_grid_msg(ref_grid)
        cm.feed_planner(pl)
        wp, (wx, wy) = worlds.next_waypoint(route, wp, x, y)
        c, s = np.cos(yaw), np.sin(yaw)
        q = make_query(c * (wx - x) + s * (wy - y), -s * (wx - x) + c * (wy - y), start_speed=v)
        want = ref_planner.plan(q)
        got = pl.plan(q)
        assert (got.status == PP_PLAN_FOUND) == (want.status == 0)
        assert (got.n_open, got.n_closed) == (want.n_open, want.n_closed)
        np.testing.assert_allclose(got.path_xy, want.path_xy, rtol=0, atol=1e-9)
        np.testing.assert_array_equal(got.speeds, want.speeds)
        np.testing.assert_allclose(got.yaw_rates, want.yaw_rates, rtol=0, atol=1e-7)
        rejected += got.n_collisions
        if want.status == 0 and len(want.speeds):
            found += 1
            v = float(want.speeds[0])
            yaw += float(want.yaw_rates[0]) * 0.05
            x += v * 0.05 * np.cos(yaw)
            y += v * 0.05 * np.sin(yaw)
    assert found >= 5 and rejected > 0, (found, rejected)
    for o in (ref_costmap, ref_planner, cm, pl):
        o.close()


def test_costmap_parameter_variants_bit_exact(oracle, cuda_device):
    """Other resolutions, non-square windows, buffers, ranges, inflation radii, global-map sizes."""
    from autonomouscar_b200 import LocalCostmap, worlds
    from autonomouscar_b200.costmap_capi import default_costmap_params
    cases = [(0.25, 300.0, 260.0, 0.075, 2.0, 5.0, False), (0.25, 260.0, 300.0, 0.2, 1.0, 8.0, True),
             (0.5, 150.0, 150.0, 0.075, 2.0, 5.0, True), (0.2, 400.0, 400.0, 0.0, 3.5, 2.0, True),
             (0.125, 520.0, 500.0, 0.075, 2.0, 10.0, False), (0.1, 1024.0, 1024.0, 0.075, 2.0, 5.0, True)]
    for k, (res, hc, wc, zb, orange, infl, with_global) in enumerate(cases):
        G = 1200
        p = default_costmap_params(res=res, height_cells=hc, width_cells=wc, z_ground_buffer=zb, obstacle_range=orange,
                                   max_inflation_radius=infl, global_height_cells=float(G), global_width_cells=float(G))
        car = (20.0 + 7 * k, -15.0 + 3 * k, 0.3 * k - 0.5)
        sc = worlds.costmap_scene(500 + k, max_reach=28.0)
        _fill_tf(sc, car)
        g = worlds.road_like_global_grid(G, 11 * k) if with_global else None
        cm = LocalCostmap(p, device=0)
        cm.set_global_grid(g)
        got = cm.build(sc)
        want, _ = oracle.costmap_build(p, sc.inputs, g, math_mode=oracle.MATH_DET)
        np.testing.assert_array_equal(got, want)
        assert (got == 100).sum() > 200
        # the batched path with these parameters: two more scenes next to this one, masks when the window allows
        sc2, sc3 = worlds.costmap_scene(600 + k, max_reach=20.0), worlds.corridor_scene(700 + k, half_width=4.0)
        _fill_tf(sc2, car)
        _fill_tf(sc3, car)
        square_enough = hc >= wc
        grids, masks = cm.build_batch([sc, sc2, sc3], want_masks=square_enough)
        np.testing.assert_array_equal(grids[0], want)
        for j, s_ in enumerate((sc2, sc3)):
            w_j, _ = oracle.costmap_build(p, s_.inputs, g, math_mode=oracle.MATH_DET)
            np.testing.assert_array_equal(grids[j + 1], w_j)
            if square_enough:
                from oracle import navigator_ref
                assert int(masks[j + 1]) == navigator_ref.region_mask(w_j, int(wc))
        cm.close()


def test_costmap_batch_equals_single_builds(oracle, cuda_device):
    """pc_build_batch: n scenes (different sensor data, car poses, sizes of scan / cloud, with the shared global
    map) through one set of launches == n pc_build calls, byte for byte, region masks included; the first
    grids are also checked against the oracle directly."""
    from autonomouscar_b200 import LocalCostmap, worlds
    from oracle import navigator_ref
    p = mcg.params()
    rng = np.random.default_rng(12)
    for with_global, n in ((True, 37), (False, 5), (True, 1), (False, 130)):
        cm = LocalCostmap(p, device=0)
        g = mcg.global_grid(p, (10.0, -20.0, 0.3), 40) if with_global else None
        if g is not None:
            cm.set_global_grid(g)
        scenes = []
        for k in range(n):
            kind = k % 3
            if kind == 0:
                sc = worlds.costmap_scene(500 + k, n_planar=int(rng.choice([640, 333])), per_ring=int(rng.choice([512, 100])))
            elif kind == 1:
                sc = worlds.corridor_scene(600 + k, half_width=float(rng.uniform(3, 10)), gap_at=(8.0 + k % 17, 11.0 + k % 17))
            else:
                sc = worlds.costmap_scene(700 + k, max_reach=float(rng.uniform(12, 29)))
            _fill_tf(sc, (float(rng.uniform(-60, 60)), float(rng.uniform(-60, 60)), float(rng.uniform(-3.1, 3.1))))
            scenes.append(sc)
        grids, masks = cm.build_batch(scenes)
        assert grids.shape == (n, cm.n_cells) and masks.shape == (n,)
        ptr, nb = cm.batch_grids_dev()
        assert nb == n and ptr
        for k, sc in enumerate(scenes):
            single = cm.build(sc)
            np.testing.assert_array_equal(grids[k], single, err_msg=f"scene {k}")
            assert int(masks[k]) == cm.front_regions() == navigator_ref.region_mask(single, 300), k
            if k < 3:
                want, _ = oracle.costmap_build(p, sc.inputs, g, math_mode=oracle.MATH_DET)
                np.testing.assert_array_equal(grids[k], want)
        assert len({int(m) & 0xF for m in masks}) >= (2 if n > 20 else 1)
        # a second, smaller batch on the same handle; device-only outputs
        g2, m2 = cm.build_batch(scenes[: max(1, n // 2)], want_host=False, want_masks=True)
        assert g2 is None
        np.testing.assert_array_equal(m2, masks[: max(1, n // 2)])
        assert cm.build_batch([], want_host=False, want_masks=False) == (None, None)
        cm.close()
