"""GPU tests of the single-process multi-device front end (pp_group_*): sharded results,
gathered over NCCL, must be byte-identical to the one-GPU run.  Uses every visible GPU
(1 on the default box: the all-gather then has a single rank)."""
import numpy as np
import pytest
import torch

from autonomouscar_b200 import PP_COLLISION_ORIENTED, default_params, worlds

pytestmark = pytest.mark.gpu


def test_group_collide_and_plan_match_single_gpu(oracle, cuda_device):
    from autonomouscar_b200 import Planner, PlannerGroup
    n_dev = torch.cuda.device_count()
    p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=5000)
    grid = worlds.block_grid(300, 300, frac=0.01, block=4, seed=9)
    worlds.clear_disc(grid, p, 0.0, 0.0, 4.0)
    grp = PlannerGroup(p, devices=list(range(n_dev)))
    assert grp.size() == n_dev
    grp.set_grid(grid)
    poses = worlds.random_poses(p, 100003, seed=8, inset_m=-3.0)   # not divisible by the rank count
    single = Planner(p, device=0)
    single.set_grid(grid)
    want = single.collide(poses)
    np.testing.assert_array_equal(grp.collide(poses), want)
    np.testing.assert_array_equal(want, oracle.collide_batch(p, grid, poses)[0])
    qs = worlds.forward_queries(37, seed=5)
    a = grp.plan_batch(qs)
    b = single.plan_batch(qs)
    for x, y in zip(a, b):
        assert x.summary() == y.summary()
        np.testing.assert_array_equal(x.path_xy, y.path_xy)
    single.close()
    grp.close()
