"""Pins the oracle -- and through it the CUDA path -- against the REFERENCE ITSELF.

oracle/_ref holds the reference's own local_planner.cpp compiled unmodified against ROS test
doubles (oracle/Makefile).  tests/golden/ref_planner.npz are outputs of that binary
(tests/golden/make_ref_golden.py).  Three layers:
  1. the oracle restatement (glibc math, like the reference binary) must reproduce the
     committed reference outputs bit for bit                        -- runs everywhere;
  2. where oracle/_ref is present, the oracle is compared live with the reference node on
     randomised queries, grids and caps, and the fixture is re-derived   -- CPU;
  3. the CUDA path (deterministic sin/cos/atan2) must give the same
     search -- same list sizes, closing order, path and profile -- with node coordinates
     within 1e-9 m of the reference's                                -- GPU.
"""
import os
import sys

import numpy as np
import pytest

from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_COLLISION_REF_AABB, PP_GRID_OCCUPANCY_MSG, PP_PLAN_CAP,
                                PP_PLAN_FOUND, make_query)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_ref_golden as mrg  # noqa: E402

MODES = (("shipped", PP_COLLISION_AS_SHIPPED), ("aabb", PP_COLLISION_REF_AABB))


def _gold():
    return np.load(os.path.join(GOLD, "ref_planner.npz"))


def _ref():
    from oracle import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return ref_binding


def _lists(res):
    from oracle.ref_binding import oracle_lists
    return oracle_lists(res)


def _check_exact(res, z, pre):
    """`res` (oracle PlanResult, glibc math) against the reference's output, bit for bit."""
    oo, oc = _lists(res)
    status = 0 if res.status == 0 else 1
    np.testing.assert_array_equal(
        np.array([status, res.n_open, res.n_closed, res.n_path, res.n_profile]), z[pre + "summary"])
    np.testing.assert_array_equal(mrg.sha(oo[:, :7]), z[pre + "open_sha"])
    np.testing.assert_array_equal(mrg.sha(oo[:-1, 7] if status == 0 else oo[:, 7]), z[pre + "open_cost_sha"])
    np.testing.assert_array_equal(oc, z[pre + "closed_state"])
    np.testing.assert_array_equal(res.path_xy, z[pre + "path_xy"])
    np.testing.assert_array_equal(res.yaw_rates, z[pre + "yaw_rates"])
    np.testing.assert_array_equal(res.durations, z[pre + "durations"])
    np.testing.assert_array_equal(res.speeds, z[pre + "speeds"])


def test_oracle_reproduces_reference_outputs(oracle):
    z = _gold()
    msg = mrg.msg_grid()
    cs = oracle.cspace_from_msg(msg, 300, 300)
    np.testing.assert_array_equal(mrg.sha(cs), z["cspace_sha"])          # LP:425-441
    for tag, mode in MODES:
        for k, t in enumerate(mrg.QUERIES):
            p = mrg.params(mode, t[4], t[5])
            res = oracle.plan(p, cs, mrg.query(t), math_mode=oracle.MATH_LIBM, want_visited=True,
                              cap_nodes=1 << 17, cap_expansions=1 << 17)
            _check_exact(res, z, f"{tag}{k}_")
            np.testing.assert_array_equal(np.packbits(res.visited), z[f"{tag}{k}_visited"])   # LP:318-326
            if k == 3:
                np.testing.assert_array_equal(_lists(res)[0][:-1], z[f"{tag}{k}_open_state"][:-1])


def test_oracle_reproduces_reference_collision_and_lattice(oracle):
    z = _gold()
    p = mrg.params(PP_COLLISION_REF_AABB, 10, 10)
    cs = oracle.cspace_from_msg(mrg.msg_grid(), 300, 300)
    want = z["collision_flags"]
    got, _ = oracle.collide_batch(p, cs, mrg.collision_poses())
    defined = want < 2                # 2 = the reference indexes out of bounds there (LP:479-493)
    assert defined.sum() > 7000 and want[defined].sum() > 500
    np.testing.assert_array_equal(got[defined], want[defined])
    q = make_query(20.0, 0.0)
    for k, v in enumerate(z["vel_in"]):
        np.testing.assert_array_equal(oracle.possible_velocities(p, q, float(v)), z[f"vel{k}"])
    yaws = [oracle.possible_yaws(p, float(h)) for h in z["yaw_in"]]
    np.testing.assert_array_equal(np.array([len(y) for y in yaws]), z["yaw_count"])
    assert set(np.unique(z["yaw_count"])) == {26, 27}                     # the LP:225-226 trap
    np.testing.assert_array_equal(mrg.sha(np.concatenate(yaws)), z["yaw_sha"])


def test_reference_binary_still_gives_the_fixture():
    ref = _ref()
    z = _gold()
    msg = mrg.msg_grid()
    for tag, variant, mode in (("shipped", ref.SHIPPED, PP_COLLISION_AS_SHIPPED), ("aabb", ref.AABB, PP_COLLISION_REF_AABB)):
        for k in (0, 5, 7):
            t = mrg.QUERIES[k]
            node = ref.RefNode(mrg.params(mode, t[4], t[5]), variant)
            node.set_grid_msg(msg)
            r = node.plan(mrg.query(t))
            np.testing.assert_array_equal(r.closed_state, z[f"{tag}{k}_closed_state"])
            np.testing.assert_array_equal(r.path_xy, z[f"{tag}{k}_path_xy"])
            node.close()


def test_oracle_matches_live_reference_on_random_queries(oracle):
    """Randomised goals, speeds, grids and caps: every list entry, path pose and profile entry
    of the reference node equals the oracle's, bit for bit."""
    ref = _ref()
    from autonomouscar_b200 import worlds
    rng = np.random.default_rng(20181212)
    n_found = n_cap = 0
    for trial in range(16):
        aabb = trial % 2 == 1
        mode = PP_COLLISION_REF_AABB if aabb else PP_COLLISION_AS_SHIPPED
        cs0 = worlds.block_grid(300, 300, frac=float(rng.uniform(0.002, 0.02)), block=int(rng.integers(2, 7)),
                                seed=int(rng.integers(1, 1 << 30)))
        max_exp = int(rng.choice([100000, 100000, 60, 25]))
        max_nodes = int(rng.choice([20000, 20000, 3000]))
        p = mrg.params(mode, max_exp, max_nodes)
        worlds.clear_disc(cs0, p, 0.0, 0.0, 4.0)
        msg = np.ascontiguousarray(cs0.reshape(-1)[::-1])     # inverse of LP:439
        q = make_query(float(rng.uniform(6, 26)), float(rng.uniform(-7, 7)),
                       goal_speed=float(rng.choice([1.5, 1.5, 1.0, 0.0])), start_speed=float(rng.uniform(0, 1.5)))
        cs = oracle.cspace_from_msg(msg, 300, 300)
        np.testing.assert_array_equal(cs, cs0)
        o = oracle.plan(p, cs, q, math_mode=oracle.MATH_LIBM, want_visited=True, cap_nodes=1 << 17,
                        cap_expansions=1 << 17)
        if o.status not in (PP_PLAN_FOUND, PP_PLAN_CAP):
            continue      # empty frontier / broken parent walk: undefined behaviour in the reference
        node = ref.RefNode(p, ref.AABB if aabb else ref.SHIPPED)
        node.set_grid_msg(msg)
        np.testing.assert_array_equal(node.cspace(), cs)
        r = node.plan(q, want_visited=True)
        oo, oc = _lists(o)
        assert (r.status == 0) == (o.status == 0), (trial, r.status, o.status)
        assert (r.n_open, r.n_closed) == (o.n_open, o.n_closed), trial
        np.testing.assert_array_equal(r.open_state[:, :7], oo[:, :7])
        cut = -1 if r.status == 0 else None   # the goal node's cost is uninitialised in the reference
        np.testing.assert_array_equal(r.open_state[:cut, 7], oo[:cut, 7])
        np.testing.assert_array_equal(r.closed_state, oc)
        np.testing.assert_array_equal(r.path_xy, o.path_xy)
        np.testing.assert_array_equal(r.yaw_rates, o.yaw_rates)
        np.testing.assert_array_equal(r.speeds, o.speeds)
        np.testing.assert_array_equal(r.durations, o.durations)
        np.testing.assert_array_equal(r.visited, o.visited)
        if r.status == 0:
            assert r.goal == (q.goal_x, q.goal_y) and (r.mp_max_speed, r.mp_max_accel) == (q.max_speed, q.max_accel)
            n_found += 1
        else:
            n_cap += 1
        node.close()
    assert n_found >= 5 and n_cap >= 2, (n_found, n_cap)


def test_oracle_matches_live_reference_collision(oracle):
    ref = _ref()
    from autonomouscar_b200 import default_params, worlds
    rng = np.random.default_rng(5)
    for res, w, h in ((0.25, 300, 300), (0.1, 640, 512)):
        p = default_params(costmap_res=res, costmap_width=w, costmap_height=h, collision_mode=PP_COLLISION_REF_AABB)
        msg = worlds.block_grid(w, h, frac=0.02, block=6, seed=11).reshape(-1)
        node = ref.RefNode(p, ref.AABB)
        node.set_grid_msg(msg)
        cs = oracle.cspace_from_msg(msg, w, h)
        ext_x, ext_y = h * res / 2 + 3, w * res / 2 + 3
        n = 30000
        poses = np.stack([rng.uniform(-ext_x, ext_x, n), rng.uniform(-ext_y, ext_y, n),
                          rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
        want, _ = node.check_collision(poses)
        got, _ = oracle.collide_batch(p, cs, poses)
        defined = want < 2
        assert defined.sum() > 20000 and want[defined].sum() > 1000
        np.testing.assert_array_equal(got[defined], want[defined])
        node.close()


@pytest.mark.gpu
def test_cuda_search_equals_the_reference_search(cuda_device):
    """The CUDA planner against the committed outputs of the reference binary: identical list
    sizes, closing (= expansion) order, path and profile; coordinates within 1e-9 m (the only
    difference is the last-ulp behaviour of sin/cos/atan2: deterministic on the GPU, glibc in
    the reference)."""
    from autonomouscar_b200 import Planner
    z = _gold()
    msg = mrg.msg_grid()
    for tag, mode in MODES:
        for k, t in enumerate(mrg.QUERIES):
            pl = Planner(mrg.params(mode, t[4], t[5]), device=0)
            pl.set_grid(msg.reshape(300, 300), layout=PP_GRID_OCCUPANCY_MSG)
            res = pl.plan(mrg.query(t), cap_nodes=1 << 17, cap_expansions=1 << 17)
            pre = f"{tag}{k}_"
            status = 0 if res.status == 0 else 1
            np.testing.assert_array_equal(
                np.array([status, res.n_open, res.n_closed, res.n_path, res.n_profile]), z[pre + "summary"])
            _, oc = _lists(res)
            want = z[pre + "closed_state"]
            np.testing.assert_allclose(oc[:, :6], want[:, :6], rtol=0, atol=1e-9)
            np.testing.assert_allclose(oc[1:, 6:], want[1:, 6:], rtol=1e-12, atol=1e-9)
            np.testing.assert_array_equal(oc[:, 5], want[:, 5])        # velocities carry no trig
            np.testing.assert_allclose(res.path_xy, z[pre + "path_xy"], rtol=0, atol=1e-9)
            np.testing.assert_allclose(res.yaw_rates, z[pre + "yaw_rates"], rtol=0, atol=1e-7)
            np.testing.assert_array_equal(res.speeds, z[pre + "speeds"])
            np.testing.assert_array_equal(res.durations, z[pre + "durations"])
            pl.close()


@pytest.mark.gpu
def test_cuda_ref_aabb_collision_equals_the_reference_booleans(cuda_device):
    """pp_collide_batch in REF_AABB mode against the booleans the REFERENCE's own checkCollision
    returned (fixture from oracle/_ref; `2` marks poses where the reference indexes out of bounds)."""
    from autonomouscar_b200 import Planner
    z = _gold()
    p = mrg.params(PP_COLLISION_REF_AABB, 10, 10)
    pl = Planner(p, device=0)
    pl.set_grid(mrg.msg_grid().reshape(300, 300), layout=PP_GRID_OCCUPANCY_MSG)
    poses = mrg.collision_poses()
    got = pl.collide(poses)
    want = z["collision_flags"]
    defined = want < 2
    np.testing.assert_array_equal(got[defined], want[defined])
    assert want[defined].sum() > 500
    pl.close()


def test_oracle_matches_live_reference_on_random_parameters(oracle):
    """The launch-file values are one point; the reference's arithmetic has parameter-dependent traps (the
    26 / 27-heading loop bound, the int() of the velocity count, search_res in the goal test).  Randomised
    node parameters, window sizes and footprints: the oracle still equals the reference node bit for bit."""
    ref = _ref()
    from autonomouscar_b200 import default_params, worlds
    rng = np.random.default_rng(424242)
    outcomes = set()
    for trial in range(14):
        aabb = trial % 2 == 0
        res = float(rng.choice([0.1, 0.2, 0.25, 0.5]))
        # windows of 40 - 60 m: no candidate can leave them (the reference's markVisited is unchecked, LP:318-326)
        W, H = int(rng.integers(40, 60) / res), int(rng.integers(40, 60) / res)
        p = default_params(
            costmap_res=res, costmap_width=W, costmap_height=H,
            collision_mode=PP_COLLISION_REF_AABB if aabb else PP_COLLISION_AS_SHIPPED,
            time_step_ms=float(rng.choice([100.0, 150.0, 200.0, 300.0])),
            velocity_res=float(rng.choice([0.05, 0.1, 0.15])),
            heading_res=float(rng.choice([0.005, 0.01, 0.0078125, 0.004])),
            heading_diff=float(rng.choice([0.03, 0.065, 0.0625, 0.1])),
            goal_pos_tolerance=float(rng.choice([0.1, 0.3])), goal_speed_tolerance=float(rng.choice([0.5, 0.2])),
            search_res=float(rng.choice([0.1, 0.25, 0.4])),
            car_length=float(rng.uniform(3.0, 5.2)), car_width=float(rng.uniform(1.2, 2.0)),
            max_nodes=int(rng.integers(400, 6000)), max_expansions=int(rng.choice([40, 300])))
        cs0 = worlds.block_grid(W, H, frac=float(rng.choice([0.0, 0.01, 0.02])), block=int(rng.choice([2, 4, 8])), seed=trial)
        worlds.clear_disc(cs0, p, 0.0, 0.0, 4.0)
        msg = np.ascontiguousarray(cs0.reshape(-1)[::-1])
        extent = min(W, H) * res / 2
        q = make_query(float(rng.uniform(2.0, min(18.0, extent - 8.0))), float(rng.uniform(-3.0, 3.0)),
                       goal_speed=float(rng.choice([0.0, 1.0, 1.5])), max_speed=float(rng.choice([1.5, 3.0])),
                       max_accel=float(rng.choice([1.0, 1.5, 2.0])), start_speed=float(rng.uniform(0.0, 1.5)))
        o = oracle.plan(p, cs0, q, math_mode=oracle.MATH_LIBM, want_visited=True, cap_nodes=1 << 15, cap_expansions=1 << 15)
        if o.status not in (PP_PLAN_FOUND, PP_PLAN_CAP):
            continue
        node = ref.RefNode(p, ref.AABB if aabb else ref.SHIPPED)
        assert node.footprint() == (p.car_length, p.car_width)
        node.set_grid_msg(msg)
        r = node.plan(q, cap_nodes=1 << 15, want_visited=True)
        oo, oc = _lists(o)
        assert (r.status == 0) == (o.status == PP_PLAN_FOUND), trial
        np.testing.assert_array_equal(r.open_state[:, :7], oo[:, :7])
        np.testing.assert_array_equal(r.closed_state, oc)
        np.testing.assert_array_equal(r.path_xy, o.path_xy)
        np.testing.assert_array_equal(r.yaw_rates, o.yaw_rates)
        np.testing.assert_array_equal(r.speeds, o.speeds)
        np.testing.assert_array_equal(r.durations, o.durations)
        np.testing.assert_array_equal(r.visited, o.visited)
        outcomes.add(o.status)
        node.close()
    assert outcomes == {PP_PLAN_FOUND, PP_PLAN_CAP}


def test_default_footprint_is_the_references_urdf():
    """setupCollision (LP:328-360) reads the footprint from the URDF through tf: car length = |dx| of
    tf(front_right_middle_sonar_link -> back_right_middle_sonar_link), car width = |dx| of
    tf(rear_right_wheel -> rear_left_wheel).  Re-derived here from the reference's own prius.urdf; the build's
    defaults (4.7 x 1.586) are those numbers, and the 2e-8 m the sonar's yaw of 1.5707 (not pi/2) takes off the
    length never moves a cell boundary of the REF_AABB box (LP:473-476)."""
    import math
    import xml.etree.ElementTree as ET
    urdf = "/root/reference/catkin_ws/src/prius_description/urdf/prius.urdf"
    if not os.path.exists(urdf):
        pytest.skip("needs /root/reference")
    joints = {}
    for j in ET.parse(urdf).getroot().iter("joint"):
        o = j.find("origin")
        if o is not None and j.find("child") is not None:
            joints[j.find("child").get("link")] = (j.find("parent").get("link"),
                                                   [float(v) for v in o.get("xyz").split()],
                                                   [float(v) for v in (o.get("rpy") or "0 0 0").split()])

    def in_frame_of(a, b):
        """origin of link b in the frame of link a (both fixed to the chassis, yaw-only rotations)"""
        (pa, xa, ra), (pb, xb, rb) = joints[a], joints[b]
        assert pa == pb == "chassis" and ra[:2] == [0.0, 0.0] and rb[:2] == [0.0, 0.0]
        dx, dy = xb[0] - xa[0], xb[1] - xa[1]
        c, s = math.cos(ra[2]), math.sin(ra[2])
        return c * dx + s * dy, -s * dx + c * dy            # R_a^T (p_b - p_a)
    from autonomouscar_b200 import default_params
    length = abs(in_frame_of("front_right_middle_sonar_link", "back_right_middle_sonar_link")[0])
    width = abs(in_frame_of("rear_right_wheel", "rear_left_wheel")[0])
    p = default_params()
    assert width == p.car_width == 1.586
    assert abs(length - p.car_length) < 1e-7 and p.car_length == 4.7
    for res in (0.25, 0.1):
        for base in range(0, 8192):
            for ext_ref, ext in ((length, p.car_length), (width, p.car_width)):
                assert int(base - ext_ref / 2 / res) == int(base - ext / 2 / res)
                assert int(base + ext_ref / 2 / res) == int(base + ext / 2 / res)


def test_library_defaults_are_the_references_launch_file():
    """pp_default_params / pc_default_params / pg_default_params against the <param> values of the reference's own
    full_sim.launch (parsed here, not copied)."""
    import xml.etree.ElementTree as ET
    launch = "/root/reference/catkin_ws/src/prius_sim/launch/full_sim.launch"
    if not os.path.exists(launch):
        pytest.skip("needs /root/reference")
    from autonomouscar_b200 import default_params
    from autonomouscar_b200.costmap_capi import default_costmap_params, default_global_costmap_params
    val = {}
    for node in ET.parse(launch).getroot().iter("node"):
        for prm in node.findall("param"):
            if prm.get("value") is not None:
                val[(node.get("name"), prm.get("name"))] = prm.get("value")
    f = lambda node, name: float(val[(node, name)])                                   # noqa: E731
    p, c, g = default_params(), default_costmap_params(), default_global_costmap_params()
    lp = "local_planner_node"
    for ours, theirs in ((p.time_step_ms, "time_step_ms"), (p.velocity_res, "velocity_res"), (p.heading_res, "heading_res"),
                         (p.heading_diff, "heading_diff"), (p.goal_pos_tolerance, "goal_pos_tolerance"),
                         (p.goal_speed_tolerance, "goal_speed_tolerance"), (p.search_res, "search_res"),
                         (p.collision_buffer_distance, "collision_buffer_distance")):
        assert ours == f(lp, theirs), theirs
    lc = "local_costmap_node"
    res = f(lc, "local_costmap_res")
    assert p.costmap_res == c.res == g.res == res
    assert p.costmap_height == c.height_cells == f(lc, "local_costmap_height") / res          # LC:20-21, LP:401-417
    assert p.costmap_width == c.width_cells == f(lc, "local_costmap_width") / res
    assert (c.z_ground_buffer, c.obstacle_range, c.max_inflation_radius) == (f(lc, "z_ground_buffer"), f(lc, "obstacle_range"),
                                                                             f(lc, "max_inflation_radius"))
    gc = "global_costmap_node"
    assert g.road_width == f(gc, "road_width") and g.num_lanes == int(f(gc, "num_lanes"))
    assert g.height_cells == c.global_height_cells == f(gc, "global_costmap_height") / res      # GC:14-15, LC:25-26
    assert g.width_cells == c.global_width_cells == f(gc, "global_costmap_width") / res
