"""Golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the
oracle): the oracle must still reproduce them (CPU), and the CUDA path must reproduce them
through the C ABI (GPU) without the oracle in the loop."""
import os
import sys

import numpy as np
import pytest

from autonomouscar_b200 import PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB, default_params, make_query

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_golden as mg  # noqa: E402


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _check_plan(r, z, pre):
    got = np.array([r.status, r.n_expansions, r.n_nodes, r.n_open, r.n_closed, r.n_path, r.n_profile,
                    r.n_candidates, r.n_collisions], dtype=np.int64)
    np.testing.assert_array_equal(got, z[pre + "summary"])
    assert np.uint64(r.tree_digest) == z[pre + "digest"][0]
    np.testing.assert_array_equal(r.path_xy, z[pre + "path_xy"])
    np.testing.assert_array_equal(r.path_node_ids, z[pre + "path_ids"])
    np.testing.assert_array_equal(r.yaw_rates, z[pre + "yaw_rates"])
    np.testing.assert_array_equal(r.speeds, z[pre + "speeds"])
    np.testing.assert_array_equal(r.expansion_ids, z[pre + "expansion_ids"])


def _plan_cases():
    p_free = default_params()
    p_obs, grid = mg.plan_world()
    for k, kw in enumerate(mg.PLAN_QUERIES):
        yield f"free{k}_", p_free, None, make_query(**kw)
        yield f"obs{k}_", p_obs, grid, make_query(**kw)


def test_oracle_reproduces_golden_plans(oracle):
    z = _load("plans.npz")
    for pre, p, g, q in _plan_cases():
        _check_plan(oracle.plan(p, g, q, math_mode=oracle.MATH_DET), z, pre)


def test_oracle_reproduces_golden_collision_sampler_rrt(oracle):
    z = _load("collision.npz")
    p, g, poses = mg.collide_world()
    np.testing.assert_array_equal(poses, z["poses"])
    for mode, tag in ((PP_COLLISION_REF_AABB, "aabb"), (PP_COLLISION_ORIENTED, "oriented")):
        p.collision_mode = mode
        np.testing.assert_array_equal(oracle.collide_batch(p, g, poses)[0], z["flags_" + tag])
    for q, want in zip(z["fp_poses"], z["fp_points"]):
        np.testing.assert_array_equal(oracle.footprint_points(p, *q), want)
    z = _load("sampler_nn.npz")
    s = oracle.sample_poses(0xC0FFEE, 3, 1 << 33, 2048, *z["box"])
    np.testing.assert_array_equal(s, z["poses"])
    z = _load("rrt.npz")
    p, g, q = mg.rrt_world()
    r = oracle.rrt_plan(p, g, q, cap_nodes=32768, cap_path=4096)
    assert np.uint64(r.tree_digest) == z["digest"][0]
    np.testing.assert_array_equal(r.path_xy, z["path_xy"])


@pytest.mark.gpu
def test_cuda_reproduces_golden_fixtures(cuda_device):
    import torch

    from autonomouscar_b200 import Planner
    z = _load("plans.npz")
    planners = {}
    for pre, p, g, q in _plan_cases():
        key = pre[:3]
        if key not in planners:
            planners[key] = Planner(p, device=0)
            if g is not None:
                planners[key].set_grid(g)
        _check_plan(planners[key].plan(q), z, pre)
    for pl in planners.values():
        pl.close()
    z = _load("collision.npz")
    p, g, poses = mg.collide_world()
    pl = Planner(p, device=0)
    pl.set_grid(g)
    for mode, tag in ((PP_COLLISION_REF_AABB, "aabb"), (PP_COLLISION_ORIENTED, "oriented")):
        pl.set_collision_mode(mode)
        np.testing.assert_array_equal(pl.collide(z["poses"]), z["flags_" + tag])
    for q, want in zip(z["fp_poses"], z["fp_points"]):
        np.testing.assert_array_equal(pl.footprint_points(*q), want)
    z = _load("sampler_nn.npz")
    box = [float(v) for v in z["box"]]
    s = pl.sample_poses_dev(0xC0FFEE, 3, 1 << 33, 2048, *box)
    pl.sync()
    np.testing.assert_array_equal(s.cpu().numpy(), z["poses"])
    idx, d2 = pl.nearest_dev(s[:1500, :2].contiguous(), s[1500:, :2].contiguous())
    pl.sync()
    np.testing.assert_array_equal(idx.cpu().numpy(), z["nn_idx"])
    np.testing.assert_array_equal(d2.cpu().numpy(), z["nn_d2"])
    pl.close()
    z = _load("rrt.npz")
    p, g, q = mg.rrt_world()
    pl = Planner(p, device=0)
    pl.set_grid(g)
    r = pl.rrt_plan(q, cap_nodes=32768, cap_path=4096)
    got = np.array([r.status, r.n_expansions, r.n_nodes, r.n_path, r.n_candidates, r.n_collisions], dtype=np.int64)
    np.testing.assert_array_equal(got, z["summary"])
    assert np.uint64(r.tree_digest) == z["digest"][0]
    np.testing.assert_array_equal(r.path_xy, z["path_xy"])
    np.testing.assert_array_equal(r.path_node_ids, z["path_ids"])
    np.testing.assert_array_equal(r.node_parent, z["parents"])
    pl.close()
