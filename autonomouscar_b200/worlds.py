"""Synthetic occupancy grids and query sets for tests and bench.py (host-side, numpy).

The reference world (prius_sim/worlds/irving.world) contains no obstacle models
(SURVEY.md F3), so BASELINE.json's synthetic grids are defined here:

  block_grid   "1 % random obstacles" as 1 % of the AREA in axis-aligned 16x16-cell blocks
               (obstacle model); free cells carry road-cost noise
               0..99 and a sprinkle of -1 (unknown) so that only `== 100` counts, as in
               LP:451.
  cell_noise_grid  the literal per-cell reading (each cell occupied with probability p).
"""
import numpy as np

from .capi import make_query


def block_grid(width, height, frac=0.01, block=16, seed=1, noise=True):
    """int8 [width, height] in the planner's m_c_space order."""
    rng = np.random.Generator(np.random.Philox(seed))
    g = np.zeros((width, height), dtype=np.int8)
    if noise:
        g[:] = rng.integers(0, 100, size=g.shape, dtype=np.int8)
        unknown = rng.random(g.shape) < 0.02
        g[unknown] = -1
    n_blocks = int(round(frac * width * height / (block * block)))
    bi = rng.integers(0, max(1, width - block), size=n_blocks)
    bj = rng.integers(0, max(1, height - block), size=n_blocks)
    for i, j in zip(bi, bj):
        g[i:i + block, j:j + block] = 100
    return g


def cell_noise_grid(width, height, p=0.01, seed=1):
    rng = np.random.Generator(np.random.Philox(seed))
    g = np.zeros((width, height), dtype=np.int8)
    g[rng.random(g.shape) < p] = 100
    return g


def clear_disc(grid, params, x, y, radius_m):
    """Free the cells within radius_m of world point (x, y) (start / goal clearings)."""
    W, H, res = params.costmap_width, params.costmap_height, params.costmap_res
    ci, cj = W // 2 - y / res, H // 2 - x / res
    r = int(np.ceil(radius_m / res)) + 1
    i0, i1 = max(0, int(ci) - r), min(W, int(ci) + r + 1)
    j0, j1 = max(0, int(cj) - r), min(H, int(cj) + r + 1)
    ii, jj = np.mgrid[i0:i1, j0:j1]
    mask = (ii - ci) ** 2 + (jj - cj) ** 2 <= (radius_m / res) ** 2
    sub = grid[i0:i1, j0:j1]
    sub[mask & (sub == 100)] = 0


def pose_box(params, inset_m=2.5):
    """World extent of the grid inset by `inset_m` (BASELINE config 3): x_lo, x_hi, y_lo, y_hi."""
    W, H, res = params.costmap_width, params.costmap_height, params.costmap_res
    x_lo, x_hi = (H // 2 - H) * res + inset_m, (H // 2) * res - inset_m
    y_lo, y_hi = (W // 2 - W) * res + inset_m, (W // 2) * res - inset_m
    return float(np.float32(x_lo)), float(np.float32(x_hi)), float(np.float32(y_lo)), float(np.float32(y_hi))


def random_poses(params, n, seed=0, inset_m=2.5):
    """Host-side pose batch (numpy RNG; the device-side Philox sampler is pp_sample_poses_dev)."""
    rng = np.random.Generator(np.random.Philox(seed))
    x_lo, x_hi, y_lo, y_hi = pose_box(params, inset_m)
    p = np.empty((n, 3), dtype=np.float32)
    p[:, 0] = rng.uniform(x_lo, x_hi, n)
    p[:, 1] = rng.uniform(y_lo, y_hi, n)
    p[:, 2] = rng.uniform(-np.pi, np.pi, n)
    return p


def forward_queries(n, seed=0, dist=(15.0, 25.0), lateral=3.0, speed=(0.0, 1.5)):
    """Plan requests shaped like local_navigator.py's stream: a goal >= 20 m ahead in the
    vehicle frame (local_navigator.py:521-527), speed/limits 1.5."""
    rng = np.random.Generator(np.random.Philox(seed))
    qs = []
    for k in range(n):
        gx = rng.uniform(*dist)
        gy = rng.uniform(-lateral, lateral)
        v0 = rng.uniform(*speed)
        qs.append(make_query(gx, gy, start_speed=v0, seed=seed, query_id=k))
    return qs
