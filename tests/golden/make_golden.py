"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).

The reference ships no golden vectors and cannot be run here (needs ROS), so these
fixtures are produced by the oracle restatement; they freeze its behaviour (any later
change to oracle/ or to the CUDA path that alters a result shows up as a diff against
these files) and they travel to the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from autonomouscar_b200 import (PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB, default_params, make_query,  # noqa: E402
                                worlds)
from oracle import binding as orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

PLAN_QUERIES = [
    dict(goal_x=20.0, goal_y=0.0, start_speed=0.0),
    dict(goal_x=20.0, goal_y=3.0, start_speed=1.5),
    dict(goal_x=12.0, goal_y=-4.0, start_speed=0.7),
]


def plan_world():
    p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=8000)
    grid = worlds.block_grid(300, 300, frac=0.01, block=4, seed=9)
    worlds.clear_disc(grid, p, 0.0, 0.0, 4.0)
    return p, grid


def collide_world():
    p = default_params(costmap_res=0.1, costmap_width=640, costmap_height=512, collision_mode=PP_COLLISION_ORIENTED)
    grid = worlds.block_grid(640, 512, frac=0.03, block=8, seed=77)
    poses = worlds.random_poses(p, 4096, seed=5, inset_m=-4.0)
    return p, grid, poses


def rrt_world():
    p = default_params(costmap_res=0.1, costmap_width=512, costmap_height=512, collision_mode=PP_COLLISION_ORIENTED,
                       time_step_ms=1000.0, rrt_samples_per_round=64, rrt_max_samples=20000, max_nodes=20000,
                       rrt_goal_tolerance=1.5)
    grid = worlds.block_grid(512, 512, frac=0.01, block=16, seed=3, noise=False)
    worlds.clear_disc(grid, p, -15.0, -15.0, 5.0)
    worlds.clear_disc(grid, p, 14.0, 12.0, 5.0)
    q = make_query(14.0, 12.0, start_x=-15.0, start_y=-15.0, start_heading=0.5, start_speed=1.0, seed=2018, query_id=1)
    return p, grid, q


def main():
    # 1. plans (free space and with obstacles)
    p_free = default_params()
    p_obs, grid = plan_world()
    blob = {}
    for k, kw in enumerate(PLAN_QUERIES):
        for tag, (p, g) in (("free", (p_free, None)), ("obs", (p_obs, grid))):
            r = orc.plan(p, g, make_query(**kw), math_mode=orc.MATH_DET)
            pre = f"{tag}{k}_"
            blob[pre + "summary"] = np.array([r.status, r.n_expansions, r.n_nodes, r.n_open, r.n_closed, r.n_path,
                                              r.n_profile, r.n_candidates, r.n_collisions], dtype=np.int64)
            blob[pre + "digest"] = np.array([r.tree_digest], dtype=np.uint64)
            blob[pre + "path_xy"] = r.path_xy.copy()
            blob[pre + "path_ids"] = r.path_node_ids.copy()
            blob[pre + "yaw_rates"] = r.yaw_rates.copy()
            blob[pre + "speeds"] = r.speeds.copy()
            blob[pre + "expansion_ids"] = r.expansion_ids.copy()
    np.savez_compressed(os.path.join(OUT, "plans.npz"), **blob)
    # 2. collision flags, both modes, + footprint lattice of three poses
    p, g, poses = collide_world()
    blob = {"poses": poses}
    for mode, tag in ((PP_COLLISION_REF_AABB, "aabb"), (PP_COLLISION_ORIENTED, "oriented")):
        p.collision_mode = mode
        blob["flags_" + tag] = orc.collide_batch(p, g, poses)[0]
    p.collision_mode = PP_COLLISION_ORIENTED
    fp_poses = np.array([[0.0, 0.0, 0.0], [12.3456, -7.25, 1.1], [-20.0, 3.0, -2.9]])
    blob["fp_poses"] = fp_poses
    blob["fp_points"] = np.stack([orc.footprint_points(p, *q) for q in fp_poses])
    np.savez_compressed(os.path.join(OUT, "collision.npz"), **blob)
    # 3. Philox sampler + nearest neighbour
    box = (-25.0, 25.0, -30.0, 20.0)
    s = orc.sample_poses(0xC0FFEE, 3, 1 << 33, 2048, *box)
    nodes = s[:1500, :2].copy()
    q = s[1500:, :2].copy()
    idx, d2 = orc.nearest(nodes, q)
    np.savez_compressed(os.path.join(OUT, "sampler_nn.npz"), box=np.array(box), poses=s, nn_idx=idx, nn_d2=d2)
    # 4. sampling planner
    p, g, q = rrt_world()
    r = orc.rrt_plan(p, g, q, cap_nodes=32768, cap_path=4096)
    np.savez_compressed(os.path.join(OUT, "rrt.npz"),
                        summary=np.array([r.status, r.n_expansions, r.n_nodes, r.n_path, r.n_candidates,
                                          r.n_collisions], dtype=np.int64),
                        digest=np.array([r.tree_digest], dtype=np.uint64), path_xy=r.path_xy.copy(),
                        path_ids=r.path_node_ids.copy(), parents=r.node_parent.copy())
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
