"""Small driver for ncu captures of the collision kernels (not a benchmark).
usage: python profiles/prof_collide.py [log2_poses] [broad 0/1] [mode 1/2] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from autonomouscar_b200 import Planner, default_params, worlds  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
broad = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
p = default_params(costmap_res=0.1, costmap_width=4096, costmap_height=4096, collision_mode=mode)
pl = Planner(p, device=0)
pl.set_grid(worlds.block_grid(4096, 4096, frac=0.01, block=16, seed=20181))
pl.set_broad_phase(bool(broad))
n = 1 << lg
poses = pl.sample_poses_dev(2018, 0, 0, n, *worlds.pose_box(p))
flags = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    pl.collide_dev(poses, out=flags)
pl.sync()
print("ms", pl.last_phase_ms()[2], "narrow", pl.last_narrow_count(), "hits", int(flags.sum()))
