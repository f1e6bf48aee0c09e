"""Randomised GPU-vs-oracle parity over parameter space (seeded): grid sizes / resolutions,
lattice parameters (26 / 27-yaw trap, 3..9 velocities), collision modes, caps, goal speeds,
sampling-planner round sizes.  Bit-exact or fail."""
import os

import numpy as np
import pytest

from autonomouscar_b200 import default_params, make_query, worlds

pytestmark = pytest.mark.gpu


def _same_plan(got, ref):
    assert got.summary() == ref.summary()
    np.testing.assert_array_equal(got.node_state.view(np.uint64), ref.node_state.view(np.uint64))
    np.testing.assert_array_equal(got.node_closed_seq, ref.node_closed_seq)
    np.testing.assert_array_equal(got.node_parent, ref.node_parent)
    np.testing.assert_array_equal(got.expansion_ids, ref.expansion_ids)
    np.testing.assert_array_equal(got.path_xy, ref.path_xy)
    np.testing.assert_array_equal(got.path_node_ids, ref.path_node_ids)
    np.testing.assert_array_equal(got.yaw_rates, ref.yaw_rates)
    np.testing.assert_array_equal(got.speeds, ref.speeds)


def test_plan_fuzz(oracle, cuda_device):
    from autonomouscar_b200 import Planner
    # PP_FUZZ_SEED / PP_FUZZ_CASES widen the sweep for one-off soak runs
    rng = np.random.default_rng(int(os.environ.get("PP_FUZZ_SEED", "20181")))
    statuses = []
    for case in range(int(os.environ.get("PP_FUZZ_CASES", "24"))):
        res = float(rng.choice([0.1, 0.2, 0.25, 0.5]))
        W, H = int(rng.integers(120, 420)), int(rng.integers(120, 420))
        p = default_params(
            costmap_res=res, costmap_width=W, costmap_height=H, collision_mode=int(rng.integers(0, 3)),
            time_step_ms=float(rng.choice([100.0, 150.0, 200.0, 300.0])),
            velocity_res=float(rng.choice([0.05, 0.1, 0.15])), heading_res=float(rng.choice([0.005, 0.01, 0.0078125])),
            heading_diff=float(rng.choice([0.03, 0.065, 0.0625, 0.1])),
            goal_pos_tolerance=float(rng.choice([0.1, 0.3])), search_res=float(rng.choice([0.1, 0.25, 0.4])),
            max_nodes=int(rng.integers(300, 5000)), max_expansions=int(rng.choice([25, 1000, 100000])))
        grid = worlds.block_grid(W, H, frac=float(rng.choice([0.0, 0.01, 0.03])), block=int(rng.choice([2, 4, 8])),
                                 seed=case)
        worlds.clear_disc(grid, p, 0.0, 0.0, 3.5)
        extent = min(W, H) * res / 2
        pl = Planner(p, device=0)
        pl.set_grid(grid)
        for _ in range(2):
            q = make_query(float(rng.uniform(2.0, min(25.0, extent))), float(rng.uniform(-4.0, 4.0)),
                           goal_speed=float(rng.choice([0.0, 1.0, 1.5])), max_speed=float(rng.choice([1.5, 3.0])),
                           max_accel=float(rng.choice([1.0, 1.5, 2.0, 3.0])), start_speed=float(rng.uniform(0.0, 1.5)))
            got = pl.plan(q, cap_nodes=8192, cap_expansions=8192)
            ref = oracle.plan(p, grid, q, math_mode=oracle.MATH_DET, cap_nodes=8192, cap_expansions=8192)
            _same_plan(got, ref)
            statuses.append(ref.status)
        pl.close()
    assert len(set(statuses)) >= 2      # found and capped cases both occurred


def test_collide_and_rrt_fuzz(oracle, cuda_device):
    from autonomouscar_b200 import Planner
    rng = np.random.default_rng(int(os.environ.get("PP_FUZZ_SEED", "77")))
    for case in range(int(os.environ.get("PP_FUZZ_CASES", "10")) * 10 // 24 or 1):
        res = float(rng.choice([0.1, 0.125, 0.2, 0.25]))
        W, H = int(rng.integers(200, 900)), int(rng.integers(200, 900))
        length, width = float(rng.uniform(3.0, 5.5)), float(rng.uniform(1.2, 2.2))
        p = default_params(costmap_res=res, costmap_width=W, costmap_height=H, collision_mode=int(rng.integers(1, 3)),
                           car_length=length, car_width=width, time_step_ms=float(rng.choice([400.0, 1000.0])),
                           search_res=float(rng.choice([0.25, 0.5])),
                           rrt_samples_per_round=int(rng.choice([7, 32, 100, 256, 700])), rrt_max_samples=4000,
                           max_nodes=3000, rrt_goal_tolerance=1.0, rrt_goal_bias=float(rng.choice([0.0, 0.05, 0.3])))
        grid = worlds.block_grid(W, H, frac=0.02, block=int(rng.choice([3, 8, 16])), seed=100 + case)
        pl = Planner(p, device=0)
        pl.set_grid(grid)
        poses = worlds.random_poses(p, 20000, seed=case, inset_m=-6.0)   # arbitrary footprints -> generic lattice path
        np.testing.assert_array_equal(pl.collide(poses), oracle.collide_batch(p, grid, poses)[0])
        ex = min(W, H) * res / 2 - 8.0
        worlds.clear_disc(grid, p, -ex * 0.5, -ex * 0.5, 4.0)
        pl.set_grid(grid)
        q = make_query(ex * 0.5, ex * 0.4, start_x=-ex * 0.5, start_y=-ex * 0.5, start_heading=float(rng.uniform(-3, 3)),
                       start_speed=1.0, seed=int(rng.integers(1, 1 << 40)), query_id=case)
        got = pl.rrt_plan(q, cap_nodes=4096, cap_path=4096)
        ref = oracle.rrt_plan(p, grid, q, cap_nodes=4096, cap_path=4096)
        assert got.summary() == ref.summary()
        np.testing.assert_array_equal(got.node_state.view(np.uint64), ref.node_state.view(np.uint64))
        np.testing.assert_array_equal(got.node_parent, ref.node_parent)
        np.testing.assert_array_equal(got.path_node_ids, ref.path_node_ids)
        pl.close()
