"""Row N4 on the GPU: the region scan of the query generator (pc_front_regions, local_navigator.py:81-257)
against the reference navigator's own flags (tests/golden/ref_navigator.npz, produced by importing the
reference module unmodified) and against the numpy restatement, then the obstacle branch inside the
costmap -> navigator -> planner chain with the costmap staying on the device."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_navigator_golden as mng  # noqa: E402
from test_navigator_ref_pin import assert_same, drive_mirror  # noqa: E402


def _gpu_mask_fn(cm):
    import torch

    def mask_of(flat, w):
        t = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.int8)).cuda()
        m = cm.front_regions_of(t.data_ptr(), w, t.numel())
        torch.cuda.synchronize()
        return m
    return mask_of


def test_region_scan_reproduces_reference_navigator(cuda_device):
    """Every costmap of the reference-navigator fixture through the CUDA scan: the masks equal the numpy
    restatement bit for bit, and the mirror driven by the GPU masks publishes exactly what the reference
    module published (path index, flags, avoidance way-points included)."""
    from autonomouscar_b200 import LocalCostmap
    from autonomouscar_b200.costmap_capi import default_costmap_params
    from oracle import navigator_ref
    cm = LocalCostmap(default_costmap_params(), device=0)
    gpu_mask = _gpu_mask_fn(cm)
    z = np.load(os.path.join(GOLD, "ref_navigator.npz"))
    n_grids, seen = 0, set()
    for seed in mng.SEEDS:
        passes = mng.passes_from_fixture(z, seed)
        for p in passes:
            if p["grid"] is not None:
                w, g = p["grid"]
                m = gpu_mask(g.reshape(-1), w)
                assert m == navigator_ref.region_mask(g.reshape(-1), w), (seed, w, hex(m))
                n_grids += 1
                seen.add(m & 0x3FFF)
        want = (z[f"s{seed}_pub"], z[f"s{seed}_state"], z[f"s{seed}_avoid"])
        assert_same(drive_mirror(mng.route(), passes, mask_of=gpu_mask), want)
    assert n_grids > 80 and len(seen) > 8
    cm.close()


def test_region_scan_random_grids_and_errors(cuda_device):
    import ctypes as C
    from autonomouscar_b200 import LocalCostmap
    from autonomouscar_b200.capi import PP_ERR_INVALID
    from autonomouscar_b200.costmap_capi import default_costmap_params
    from oracle import navigator_ref
    import torch
    cm = LocalCostmap(default_costmap_params(), device=0)
    gpu_mask = _gpu_mask_fn(cm)
    rng = np.random.default_rng(5)
    for trial in range(60):
        w = int(rng.choice([300, 300, 512, 299, 240, 210, 181, 180, 179, 161, 141, 121, 120, 64, 1]))
        g = np.zeros((w, w), dtype=np.int8)
        dense = rng.random() < 0.3
        g[rng.random((w, w)) < (0.3 if dense else 0.0005)] = 100
        g[rng.random((w, w)) < 0.01] = 99
        assert gpu_mask(g.reshape(-1), w) == navigator_ref.region_mask(g.reshape(-1), w), (trial, w)
    # a grid with fewer than width * width cells: the reference would raise IndexError (:108); we refuse
    t = torch.zeros(300 * 200, dtype=torch.int8, device="cuda")
    m = C.c_uint32()
    assert cm._L.pc_front_regions_dev(cm._h, C.c_void_p(t.data_ptr()), 300, t.numel(), C.byref(m)) == PP_ERR_INVALID
    assert cm._L.pc_front_regions_dev(cm._h, None, 300, 90000, C.byref(m)) == PP_ERR_INVALID
    cm.close()


def test_obstacle_branch_in_the_device_chain(oracle, cuda_device):
    """A box on the centre line of a wide corridor, approached in steps: costmap built on the device, region
    flags read from the device grid, the navigator's obstacle branch picks the goal (slow down -> swerve left
    -> stop), the planner plans to it on the same device grid.  Masks, goals and plans equal the CPU chain
    (oracle costmap -> numpy scan -> same navigator -> oracle planner) at every step."""
    from autonomouscar_b200 import (PP_COLLISION_ORIENTED, LocalCostmap, Planner, default_params, make_query, rollout,
                                    worlds)
    from autonomouscar_b200.costmap_capi import default_costmap_params
    from oracle import navigator_ref
    cp = default_costmap_params()
    cm = LocalCostmap(cp, device=0)
    p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=20000)
    pl = Planner(p, device=0)
    nav_gpu, nav_cpu = rollout.LocalNavigator(), rollout.LocalNavigator()
    route = [(-50.0, 0.0), (200.0, 0.0)]
    nav_gpu.on_path(route)
    nav_cpu.on_path(route)
    box_x = 30.0
    speeds, flags = [], []
    for k, x in enumerate((0.0, 4.0, 8.0, 13.0, 17.0, 21.0, 23.5)):
        d = box_x - x
        sc = worlds.corridor_scene(40 + k, half_width=9.0, gap_at=(d, d + 3.0) if d < 29 else None)
        worlds.fill_costmap_tf(sc, (x, 0.0, 0.0))
        cm.build(sc, want_host=False)
        want_msg, _ = oracle.costmap_build(cp, sc.inputs, None, math_mode=oracle.MATH_DET)
        m = cm.front_regions()
        assert m == navigator_ref.region_mask(want_msg, 300), (k, hex(m))
        cm.feed_planner(pl)
        for nav, mask in ((nav_gpu, m), (nav_cpu, navigator_ref.region_mask(want_msg, 300))):
            nav.on_pose(x, 0.0, 0.0)
            nav.on_costmap(mask)
        goal, goal_cpu = nav_gpu.spin_once(), nav_cpu.spin_once()
        assert goal == goal_cpu
        f = rollout.RegionFlags(m)
        flags.append(tuple(f.F))
        speeds.append(goal[2])
        q = make_query(goal[0] - x, goal[1], goal_speed=goal[2], max_speed=goal[3], max_accel=goal[4], start_speed=1.0)
        r = pl.plan(q)
        ref = oracle.plan(p, oracle.cspace_from_msg(want_msg, 300, 300), q, math_mode=oracle.MATH_DET)
        assert r.summary() == ref.summary()
        np.testing.assert_array_equal(r.path_xy, ref.path_xy)
    print("approach: F flags", flags, "goal speeds", speeds)
    assert speeds[0] == 1.5 and not any(flags[0])            # nothing within 37.5 m: the route way-point
    assert 0.0 in speeds and 1.0 in speeds                   # swerve and stop way-points were issued
    assert len(set(flags)) >= 3
    cm.close()
    pl.close()


def test_sweep_with_per_car_costmaps(oracle, cuda_device):
    """A scenario sweep with sensors in the loop: every tick, ALL cars' costmaps and region masks come from one
    pc_build_batch call, the navigators turn them into goals (route / slow down / swerve / stop) and one
    pp_plan_batch call plans for all cars.  The same loop on the CPU (oracle costmap -> numpy scan -> oracle
    planner) drives every car to bit-identical states."""
    from autonomouscar_b200 import PP_COLLISION_AS_SHIPPED, PP_PLAN_FOUND, LocalCostmap, Planner, default_params, rollout, worlds
    from autonomouscar_b200.costmap_capi import default_costmap_params
    from oracle import navigator_ref
    cp = default_costmap_params()
    cm = LocalCostmap(cp, device=0)
    p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000)
    pl = Planner(p, device=0)
    route = rollout.densify_route([[-50.0, 0.0], [200.0, 0.0]])
    n = 12
    starts = [(2.5 * k, 0.02 * (k % 3), 0.0, 1.0) for k in range(n)]
    box_x = [40.0 if k % 2 == 0 else None for k in range(n)]       # even cars meet a box, odd cars an open road
    gpu, cpu = rollout.RouteRollout(starts, route=route), rollout.RouteRollout(starts, route=route)

    def oracle_batch(queries):
        out = []
        for q in queries:
            r = oracle.plan(p, None, q, math_mode=oracle.MATH_DET, want_nodes=False, cap_nodes=8)
            out.append((r.status == PP_PLAN_FOUND, r.yaw_rates.copy(), r.durations.copy(), r.speeds.copy()))
        return out

    speeds_seen = set()
    for tick in range(8):
        np.testing.assert_array_equal(gpu.trace[-1] if gpu.trace else np.array(starts), cpu.trace[-1] if cpu.trace else np.array(starts))
        scenes = []
        for k in range(n):
            d = None if box_x[k] is None else box_x[k] - float(gpu.x[k])
            sc = worlds.corridor_scene(1000 + 17 * tick + k, half_width=9.0, gap_at=(d, d + 3.0) if d is not None and 3 < d < 29 else None)
            worlds.fill_costmap_tf(sc, (float(gpu.x[k]), float(gpu.y[k]), float(gpu.yaw[k])))
            scenes.append(sc)
        _, masks = cm.build_batch(scenes, want_host=False)
        cpu_masks = [navigator_ref.region_mask(oracle.costmap_build(cp, sc.inputs, None, math_mode=oracle.MATH_DET)[0], 300)
                     for sc in scenes]
        assert [int(m) for m in masks] == cpu_masks
        qs = gpu.queries(masks)
        speeds_seen |= {q.goal_speed for q in qs}
        gpu.tick(rollout.plan_batch_cuda(pl), masks)
        cpu.tick(oracle_batch, cpu_masks)
        # the cars crawl (1 m/s): push the even ones forward so that the box passes through every region
        for ro in (gpu, cpu):
            ro.x[::2] += 4.0
    np.testing.assert_array_equal(np.array(gpu.trace), np.array(cpu.trace))
    assert gpu.found > 0.5 * gpu.plans
    assert {1.5, 1.0, 0.0} <= speeds_seen, speeds_seen        # route, swerve and stop goals all occurred
    cm.close()
    pl.close()
