"""BASELINE configs[3]: dense-clutter 8192x8192 grid @0.1 m (1 % of the area in 16x16-cell
blocks), 1M-sample sampling-planner run on one GPU, start / goal at opposite corners inset
10 m.  usage: python profiles/prof_rrt_config4.py [K] [max_samples] [time_step_ms] [goal_tolerance]
(goal_tolerance 0 switches the goal box off: the run spends its whole sample budget)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autonomouscar_b200 import PP_COLLISION_ORIENTED, Planner, default_params, make_query, worlds  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 1000.0
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
N = 8192
p = default_params(costmap_res=0.1, costmap_width=N, costmap_height=N, collision_mode=PP_COLLISION_ORIENTED,
                   time_step_ms=dt, rrt_samples_per_round=K, rrt_max_samples=S, max_nodes=S + 1,
                   rrt_goal_tolerance=tol, rrt_goal_bias=0.05)
t0 = time.time()
g = worlds.block_grid(N, N, frac=0.01, block=16, seed=4, noise=False)
half = N * 0.1 / 2
sx, sy, gx, gy = -half + 10.0, -half + 10.0, half - 10.0, half - 10.0
worlds.clear_disc(g, p, sx, sy, 6.0)
worlds.clear_disc(g, p, gx, gy, 6.0)
print("grid built in %.1f s" % (time.time() - t0))
pl = Planner(p, device=0)
pl.set_grid(g)
q = make_query(gx, gy, start_x=sx, start_y=sy, start_heading=0.785, start_speed=1.0, max_speed=1.5, seed=2018, query_id=0)
for rep in range(2):
    t0 = time.perf_counter()
    r = pl.rrt_plan(q, cap_nodes=8, cap_path=65536, want_nodes=False)
    wall = time.perf_counter() - t0
    print("rep", rep, r.summary(), "wall %.3f s" % wall, "kernel ms", pl.last_phase_ms()[4],
          "samples/s %.3e" % (r.n_candidates / wall))
if r.n_path:
    print("path", r.n_path, r.path_xy[0], r.path_xy[-1])
