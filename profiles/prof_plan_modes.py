"""Plan-batch timing per collision mode (diagnostic): python profiles/prof_plan_modes.py [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB, PlanBatch, Planner,
                                default_params, worlds)
from autonomouscar_b200.capi import pp_query

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
grid = worlds.block_grid(300, 300, frac=0.01, block=4, seed=9)
qs = worlds.forward_queries(n, seed=11)
arr = (pp_query * n)(*qs)
for name, mode in (("shipped", PP_COLLISION_AS_SHIPPED), ("aabb", PP_COLLISION_REF_AABB), ("oriented", PP_COLLISION_ORIENTED)):
    p = default_params(collision_mode=mode, max_nodes=10000)
    g = grid.copy()
    worlds.clear_disc(g, p, 0.0, 0.0, 4.0)
    pl = Planner(p, device=0)
    pl.set_grid(g)
    buf = PlanBatch(n, cap_path=256)
    pl.plan_batch_into(arr, buf)
    ts, ks = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        pl.plan_batch_into(arr, buf)
        ts.append(time.perf_counter() - t0)
        ks.append(pl.last_phase_ms()[4])
    rec = buf.records()
    print(f"{name:9s} n={n} wall {1e3 * min(ts):8.2f} ms kernel {min(ks):8.2f} ms  found {(rec[:, 0] == 0).sum()} "
          f"mean nodes {rec[:, 2].mean():.0f} max nodes {rec[:, 2].max()} mean exp {rec[:, 1].mean():.0f} max exp {rec[:, 1].max()}")
    pl.close()
