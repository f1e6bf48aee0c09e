"""Batched local costmap builder (pc_build_batch): python profiles/prof_costmap_batch.py [n_scenes] [global_cells] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autonomouscar_b200 import LocalCostmap, worlds
from autonomouscar_b200.costmap_capi import default_costmap_params

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
G = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
p = default_costmap_params(global_height_cells=float(G), global_width_cells=float(G))
car = (30.0, -12.0, 0.7)
cm = LocalCostmap(p, device=0)
if G > 0:
    cm.set_global_grid(worlds.road_like_global_grid(G, 17, zero_at=worlds.costmap_corner_global_cells(p, *car)))
base = [worlds.costmap_scene(s) for s in range(8)]
for sc in base:
    worlds.fill_costmap_tf(sc, car)
scenes = [base[k % 8] for k in range(N)]
cm.build_batch(scenes, want_host=False)
t0 = time.perf_counter()
for _ in range(reps):
    cm.build_batch(scenes, want_host=False)
dt = (time.perf_counter() - t0) / reps
print(f"{N} scenes: {1e3 * dt:.2f} ms per pc_build_batch call (masks only), device {cm.last_build_ms():.2f} ms "
      f"=> {N / dt:.0f} grids/s, {1e3 * cm.last_build_ms() / N:.1f} us of device time per grid")
cm.close()
