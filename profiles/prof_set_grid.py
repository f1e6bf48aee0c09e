"""Time of pp_set_grid_dev (all derived maps of one costmap, device-resident input): python profiles/prof_set_grid.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autonomouscar_b200 import PP_COLLISION_ORIENTED, Planner, default_params, worlds

for res, W, H in ((0.25, 300, 300), (0.1, 333, 517), (0.1, 4096, 4096), (0.1, 8192, 8192)):
    p = default_params(costmap_res=res, costmap_width=W, costmap_height=H, collision_mode=PP_COLLISION_ORIENTED)
    g = torch.from_numpy(worlds.block_grid(W, H, frac=0.01, block=16, seed=1)).cuda()
    pl = Planner(p, device=0)
    pl.set_grid(g)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        pl.set_grid(g)          # (synchronous: returns when every derived map is built)
    dt = (time.perf_counter() - t0) / 10
    print(f"{W}x{H} @ {res}: pp_set_grid_dev {dt * 1e3:.3f} ms")
    pl.close()
