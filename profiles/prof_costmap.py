"""Local costmap builder timing (row N2): python profiles/prof_costmap.py [global_cells] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from autonomouscar_b200 import LocalCostmap, worlds
from autonomouscar_b200.costmap_capi import default_costmap_params
from test_gpu_costmap import _fill_tf

G = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
p = default_costmap_params(global_height_cells=float(G), global_width_cells=float(G))
car = (30.0, -12.0, 0.7)
cm = LocalCostmap(p, device=0)
cm.set_global_grid(worlds.road_like_global_grid(G, 17, zero_at=worlds.costmap_corner_global_cells(p, *car)))
scenes = [worlds.costmap_scene(s) for s in range(4)]
for sc in scenes:
    _fill_tf(sc, car)
for want_host in (True, False):
    for sc in scenes:
        cm.build(sc, want_host=want_host)
    ms = []
    t0 = time.perf_counter()
    for k in range(reps):
        cm.build(scenes[k % 4], want_host=want_host)
        ms.append(cm.last_build_ms())
    dt = (time.perf_counter() - t0) / reps
    print(f"want_host={want_host}: {1e6 * dt:.1f} us per pc_build call, device (H2D + graph) {1e3 * np.median(ms):.1f} us "
          f"=> {1 / dt:.0f} grids/s")
cm.close()
