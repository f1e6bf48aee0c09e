"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autonomouscar_b200 import (PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB, PP_GRID_OCCUPANCY_MSG, Planner,  # noqa: E402
                                default_params, make_query, worlds)

for res, W, H in ((0.1, 333, 517), (0.25, 300, 300)):
    p = default_params(costmap_res=res, costmap_width=W, costmap_height=H, collision_mode=PP_COLLISION_ORIENTED,
                       max_nodes=1500, rrt_samples_per_round=50, rrt_max_samples=1500, time_step_ms=400.0 if res < 0.2 else 150.0)
    g = worlds.block_grid(W, H, frac=0.02, block=6, seed=3)
    worlds.clear_disc(g, p, 0.0, 0.0, 4.0)
    pl = Planner(p, device=0)
    pl.set_grid(g)
    pl.set_grid(np.ascontiguousarray(g.ravel()[::-1]), PP_GRID_OCCUPANCY_MSG)
    poses = worlds.random_poses(p, 5000, seed=1, inset_m=-6.0)
    for mode in (PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED):
        pl.set_collision_mode(mode)
        pl.collide(poses)
        for broad in (0, 1):
            pl.set_broad_phase(broad)
            pl.collide_dev(torch.from_numpy(poses).cuda())
        e = np.concatenate([poses[:, :2], poses[:, :2] + 1.0, poses[:, 2:3]], 1).astype(np.float32)
        pl.collide_edges_dev(torch.from_numpy(e).cuda())
    pl.sync()
    # round 2: the TMA-staged REF_AABB stream kernel (>= 8192 aligned poses + a ragged tail), bit-packed flags, the
    # 6-byte host format, device-resident plan records, the cooperative sampling-planner kernel with a regrouped tree
    from autonomouscar_b200.capi import pp_query, quantise_poses
    big = worlds.random_poses(p, 20011, seed=2, inset_m=-6.0)
    pl.set_collision_mode(PP_COLLISION_REF_AABB)
    f = pl.collide_dev(torch.from_numpy(big).cuda())
    pl.unpack_flags_dev(pl.pack_flags_dev(f), f.numel())
    for m in (PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED):          # bits written by the kernel / packed afterwards
        pl.set_collision_mode(m)
        pl.collide_bits_dev(torch.from_numpy(big).cuda())
    pl.set_collision_mode(PP_COLLISION_REF_AABB)
    q16, quant, _ = quantise_poses(big, *worlds.pose_box(p, inset_m=-6.0))
    pl.collide_q16(q16, quant)
    pl.set_collision_mode(PP_COLLISION_ORIENTED)
    qs = [make_query(5.0 + k * 0.1, 0.3, start_speed=1.0) for k in range(9)]
    pl.plan_batch_dev((pp_query * 9)(*qs), 9, cap_path=40)
    pl.sync()
    pl.footprint_points(1.0, 2.0, 0.3)
    s = pl.sample_poses_dev(1, 2, 3, 4097, -5, 5, -5, 5)
    pl.nearest_dev(s[:3000, :2].contiguous(), s[3000:, :2].contiguous())
    keys = torch.randn(1000, 6, dtype=torch.float64, device="cuda")
    pl.find_exact_dev(keys, keys[::7].contiguous())
    pl.plan(make_query(6.0, 0.5, start_speed=1.0))
    pl.plan_batch([make_query(5.0 + k * 0.1, 0.3, start_speed=1.0) for k in range(5)], want_nodes=True)
    pl.rrt_plan(make_query(8.0, 3.0, start_speed=1.0, seed=5))
    pl.rrt_plan_batch([make_query(8.0, 3.0, start_speed=1.0, seed=5, query_id=k) for k in range(3)])
    pl.sync()
    pl.close()
# the persistent kernel past its first regroup (> 8192 nodes): open grid, many samples
p = default_params(costmap_res=0.1, costmap_width=2048, costmap_height=2048, collision_mode=PP_COLLISION_ORIENTED,
                   time_step_ms=1000.0, rrt_samples_per_round=256, rrt_max_samples=30000, max_nodes=40000, rrt_goal_tolerance=0.0)
g = worlds.block_grid(2048, 2048, frac=0.01, block=16, seed=4, noise=False)
worlds.clear_disc(g, p, -90.0, -90.0, 6.0)
pl = Planner(p, device=0)
pl.set_grid(g)
r = pl.rrt_plan(make_query(90.0, 90.0, start_x=-90.0, start_y=-90.0, start_heading=0.785, start_speed=1.0, seed=3), cap_nodes=8,
                cap_path=64, want_nodes=False)
print("rrt nodes", r.n_nodes)
pl.close()
print("sanitize pass done")
