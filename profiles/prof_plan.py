"""Small driver for ncu captures of plan_kernel (not a benchmark).
usage: python profiles/prof_plan.py [n_queries] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autonomouscar_b200 import PP_COLLISION_ORIENTED, PlanBatch, Planner, default_params, worlds  # noqa: E402
from autonomouscar_b200.capi import pp_query  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=10000)
grid = worlds.block_grid(300, 300, frac=0.01, block=4, seed=9)
worlds.clear_disc(grid, p, 0.0, 0.0, 4.0)
pl = Planner(p, device=0)
pl.set_grid(grid)
qs = worlds.forward_queries(n, seed=11)
arr = (pp_query * n)(*qs)
buf = PlanBatch(n, cap_path=256)
for _ in range(reps):
    t0 = time.perf_counter()
    pl.plan_batch_into(arr, buf)
    dt = time.perf_counter() - t0
print("n", n, "wall ms", dt * 1e3, "kernel ms", pl.last_phase_ms()[4], "mean exp", buf.field("n_expansions").mean(),
      "mean nodes", buf.field("n_nodes").mean())
