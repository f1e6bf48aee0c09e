import sys, time, torch
sys.path.insert(0,'/root/repo')
from autonomouscar_b200 import Planner, default_params, worlds
p = default_params(costmap_res=0.1, costmap_width=4096, costmap_height=4096, collision_mode=2)
pl = Planner(p, device=0)
pl.set_grid(worlds.block_grid(4096, 4096, frac=0.01, block=16, seed=20181))
n = 1 << 24
poses = pl.sample_poses_dev(2018, 0, 0, n, *worlds.pose_box(p)); pl.sync()
h = torch.empty((n,3), dtype=torch.float32, pin_memory=True); h.copy_(poses.view(n,3))
f = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for _ in range(3): pl.collide(h, out=f)
t0=time.perf_counter()
for _ in range(20): pl.collide(h, out=f)
dt=(time.perf_counter()-t0)/20
print("e2e ms", dt*1e3, "G poses/s", n/dt/1e9, "H2D GB/s", 12*n/dt/1e9)
