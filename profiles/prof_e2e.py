"""The host-pointer collision calls on pinned buffers (float32 and int16 poses): python profiles/prof_e2e.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autonomouscar_b200 import Planner, default_params, worlds
from autonomouscar_b200.capi import quantise_poses

p = default_params(costmap_res=0.1, costmap_width=4096, costmap_height=4096, collision_mode=2)
pl = Planner(p, device=0)
pl.set_grid(worlds.block_grid(4096, 4096, frac=0.01, block=16, seed=20181))
n = 1 << 24
box = worlds.pose_box(p)
poses = pl.sample_poses_dev(2018, 0, 0, n, *box); pl.sync()
h = torch.empty((n, 3), dtype=torch.float32, pin_memory=True); h.copy_(poses.view(n, 3))
f = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for _ in range(3): pl.collide(h, out=f)
t0 = time.perf_counter()
for _ in range(20): pl.collide(h, out=f)
dt = (time.perf_counter() - t0) / 20
print("float32: e2e ms %.4f  G poses/s %.3f  H2D GB/s %.2f" % (dt * 1e3, n / dt / 1e9, 12 * n / dt / 1e9))
q_np, quant, _ = quantise_poses(h.numpy(), *box)
hq = torch.empty((n, 3), dtype=torch.int16, pin_memory=True); hq.copy_(torch.from_numpy(q_np))
for _ in range(3): pl.collide_q16(hq, quant, out=f)
t0 = time.perf_counter()
for _ in range(20): pl.collide_q16(hq, quant, out=f)
dt = (time.perf_counter() - t0) / 20
print("int16:   e2e ms %.4f  G poses/s %.3f  H2D GB/s %.2f" % (dt * 1e3, n / dt / 1e9, 6 * n / dt / 1e9))
