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


# the 12 waypoints of the reference's global route (global_path_planner.py:101-102), world frame
ROUTE_X = [26.4527, 39.5884, 52.7228, 69.0941, 83.7454, 105.948, 123.622, 142.292, 162.782, 172.996, 171.868, 171.274]
ROUTE_Y = [13.998, 14.5571, 14.5571, 15.1161, 15.6748, 16.14200, 16.142, 16.0263, 16.2719, 48.2004, 82.6559, 122.569]


def road_world(path):
    """BASELINE configs[0] world from the committed fixture (see tests/golden/make_road_world.py):
    int8 [4096, 4096] m_c_space grid in the vehicle frame of the spawn pose, off-road = 100."""
    z = np.load(path)
    n = int(z["n"][0])
    occ = np.unpackbits(z["packed"])[: n * n].reshape(n, n).astype(bool)
    g = np.zeros((n, n), dtype=np.int8)
    g[occ] = 100
    return g, float(z["res"][0]), tuple(float(v) for v in z["spawn"])


def route_queries(spawn=(0.0, 15.0), lookahead=20.0, spacing=1.0):
    """The reference's own query rule (local_navigator.py:27-46, 586-609): densify the route to
    `spacing` metres and take the first waypoint at least `lookahead` metres from the car.
    Returns the goal of the first query in the vehicle frame of the spawn pose (yaw 0)."""
    pts = [(ROUTE_X[0], ROUTE_Y[0])]
    for k in range(len(ROUTE_X) - 1):
        x0, y0, x1, y1 = ROUTE_X[k], ROUTE_Y[k], ROUTE_X[k + 1], ROUTE_Y[k + 1]
        d = float(np.hypot(x1 - x0, y1 - y0)) / spacing
        if d > 1:
            n = int(np.ceil(d))
            for m in range(1, n + 1):
                pts.append((x0 + m * (x1 - x0) / n, y0 + m * (y1 - y0) / n))
        else:
            pts.append((x1, y1))
    for x, y in pts:
        if np.hypot(x - spawn[0], y - spawn[1]) >= lookahead:
            return x - spawn[0], y - spawn[1]
    return pts[-1][0] - spawn[0], pts[-1][1] - spawn[1]


# ---------------------------------------------------------------------------------------------
# local costmap scenes (row N2).  Synthetic stand-ins for what Gazebo would publish: two planar
# 640-ray scans (prius.urdf:537-550, 569-582), one 16 x 512 block-laser cloud (prius.urdf:505-518)
# and the tf frames local_costmap.cpp looks up.  Everything stays inside the 75 m window, because
# the reference indexes its grid unchecked (see oracle/ref_costmap_harness.cpp).
# ---------------------------------------------------------------------------------------------
def costmap_tfs(car_x, car_y, car_yaw):
    """(target, source) -> (x, y, z, roll, pitch, yaw) for every lookupTransform of local_costmap.cpp."""
    c, s = np.cos(car_yaw), np.sin(car_yaw)
    inv = (-(c * car_x + s * car_y), -(-s * car_x + c * car_y))
    return {
        ("base_link", "center_laser_link"): (0.4, 0.0, 1.8, 0.0, 0.0, 0.0),
        ("front_right_laser_link", "center_laser_link"): (2.7, 1.0, 1.3, 0.05, 0.0, 0.0),
        ("front_left_laser_link", "center_laser_link"): (2.7, -1.0, 1.3, 0.05, 0.0, 0.0),
        ("map", "front_right_laser_link"): (car_x, car_y, 0.5, 0.0, 0.0, car_yaw),
        ("map", "front_left_laser_link"): (car_x, car_y, 0.5, 0.0, 0.0, car_yaw),
        ("center_laser_link", "map"): (inv[0], inv[1], -1.8, 0.0, 0.0, -car_yaw),
        ("map", "base_link"): (car_x, car_y, 0.0, 0.0, 0.0, car_yaw),
    }


def costmap_scene(seed, n_planar=640, rings=16, per_ring=512, max_reach=29.0):
    """Random but plausible sensor messages -> costmap_capi.CostmapScene (tf numbers NOT filled)."""
    from .costmap_capi import CostmapScene
    rng = np.random.default_rng(seed)
    angle_min = np.float32(-2.26889)
    angle_inc = np.float32((2.2689 + 2.26889) / (n_planar - 1))
    range_max = np.float32(30.0)

    def planar():
        r = rng.uniform(0.4, max_reach, n_planar)
        walls = rng.uniform(3.0, max_reach, 6)          # a few wall-like runs of similar range
        for w in walls:
            a = rng.integers(0, n_planar - 40)
            r[a:a + 40] = w + rng.normal(0, 0.05, 40)
        r[rng.random(n_planar) < 0.15] = 31.0           # no return: beyond range_max, skipped (LC:111)
        return np.clip(r, 0.3, 31.0).astype(np.float32)

    n = rings * per_ring
    ang = np.tile(np.linspace(-3.14, 3.14, per_ring), rings)
    rad = np.repeat(np.linspace(4.0, max_reach, rings), per_ring) + rng.normal(0, 0.05, n)
    z = np.full(n, -1.8) + rng.normal(0, 0.01, n)       # ground returns
    obs = rng.random(n) < 0.12
    z[obs] = rng.uniform(-1.5, 0.6, obs.sum())          # above the ground buffer: obstacles
    near = rng.random(n) < 0.03
    rad[near] = rng.uniform(0.3, 2.0, near.sum())       # inside obstacle_range
    z[rng.random(n) < 0.01] = 0.0                       # exercises the `!pt.z` quirk of LC:194
    cloud = np.stack([rad * np.cos(ang), rad * np.sin(ang), z], 1).astype(np.float32)
    return CostmapScene(planar(), planar(), cloud, (angle_min, angle_inc, range_max))


def road_like_global_grid(cells, offset, zero_at=None):
    """int8 [cells*cells] with values 0..99 in broad diagonal bands (what global_costmap.cpp's lane
    profile looks like from far away) and 99 elsewhere; pure integer arithmetic, so the same
    everywhere.  `zero_at` = flat indices forced to 0."""
    i = np.arange(cells, dtype=np.int64)
    band = (i[:, None] * 3 + i[None, :] * 2 + int(offset)) % 160
    g = np.where(band < 100, band, 99).astype(np.int8)
    g = g.reshape(-1)
    if zero_at is not None:
        g[np.asarray(zero_at, dtype=np.int64)] = 0
    return g


def costmap_corner_global_cells(params, car_x, car_y, car_yaw, halo=3):
    """Flat global-grid indices around the map cell under the window's first road-overlay sample
    (base_link (-H*res/2, -W*res/2)); its LOCAL index is -1 in the reference (LC:263), i.e. an
    out-of-bounds write unless the global value there is 0 -- test scenes zero these cells."""
    ux, uy = -params.height_cells * params.res / 2, -params.width_cells * params.res / 2
    c, s = np.cos(car_yaw), np.sin(car_yaw)
    gx, gy = c * ux - s * uy + car_x, s * ux + c * uy + car_y
    hg, wg = int(params.global_height_cells), int(params.global_width_cells)
    cx, cy = int(hg / 2 - gx / params.res), int(wg / 2 - gy / params.res)
    out = []
    for dx in range(-halo, halo + 1):
        for dy in range(-halo, halo + 1):
            loc = (hg - (cx + dx)) + (wg - (cy + dy)) * hg
            if 0 <= loc < hg * wg:
                out.append(loc)
    return out


def fill_costmap_tf(scene, car):
    """Fill scene.inputs' tf-derived numbers for the frames of costmap_tfs(*car) directly (planar
    car pose, lasers rolled about x only) -- what ros/local_costmap_b200_node.cpp reads from tf."""
    tfs = costmap_tfs(*car)
    for name, scan in (("front_right_laser_link", scene.inputs.right), ("front_left_laser_link", scene.inputs.left)):
        x, y, z, roll, _, _ = tfs[(name, "center_laser_link")]
        scan.roll, scan.tf_x, scan.tf_y = roll, x, y
        scan.z_map = tfs[("map", name)][2]
        scan.inv_x, scan.inv_y = -x, -(np.cos(roll) * y + np.sin(roll) * z)   # origin of the inverse: -(R^T t)
    scene.inputs.center_z = tfs[("center_laser_link", "map")][2]
    c, s = np.cos(car[2]), np.sin(car[2])
    m = [c, -s, 0, car[0], s, c, 0, car[1], 0, 0, 1, 0]
    inv = [c, s, 0, -(c * car[0] + s * car[1]), -s, c, 0, -(-s * car[0] + c * car[1]), 0, 0, 1, 0]
    for i in range(12):
        scene.inputs.map_from_base[i], scene.inputs.base_from_map[i] = m[i], inv[i]


def densified_route(spacing=1.0):
    """local_navigator.py:27-46: the global route with way-points every `spacing` metres (map frame)."""
    pts = [(ROUTE_X[0], ROUTE_Y[0])]
    for k in range(len(ROUTE_X) - 1):
        x0, y0, x1, y1 = ROUTE_X[k], ROUTE_Y[k], ROUTE_X[k + 1], ROUTE_Y[k + 1]
        d = float(np.hypot(x1 - x0, y1 - y0)) / spacing
        if d > 1:
            n = int(np.ceil(d))
            for m in range(1, n + 1):
                pts.append((x0 + m * (x1 - x0) / n, y0 + m * (y1 - y0) / n))
        else:
            pts.append((x1, y1))
    return pts


def next_waypoint(route, start_index, x, y, lookahead=20.0):
    """local_navigator.py:586-609: the first way-point from `start_index` on that is at least
    `lookahead` metres from the car.  Returns (index, (wx, wy)); the last one if none is that far."""
    for k in range(start_index, len(route)):
        if np.hypot(route[k][0] - x, route[k][1] - y) >= lookahead:
            return k, route[k]
    return len(route) - 1, route[-1]


def corridor_scene(seed, half_width=6.0, reach=29.0, rings=16, per_ring=512, gap_at=None):
    """Sensor messages of a car driving between two walls (block-laser returns at y = +-half_width, 0.8 m above
    the ground buffer) over open ground; the planar lasers see nothing (ranges beyond range_max).  `gap_at` =
    (x0, x1) removes the LEFT wall there and puts a 2 m box on the centre line instead (something to steer round)."""
    from .costmap_capi import CostmapScene
    rng = np.random.default_rng(seed)
    n = rings * per_ring
    ang = np.tile(np.linspace(-3.14, 3.14, per_ring), rings)
    rad = np.repeat(np.linspace(2.5, reach, rings), per_ring) + rng.normal(0, 0.03, n)
    gx, gy = rad * np.cos(ang), rad * np.sin(ang)
    keep = np.abs(gy) < half_width - 0.3                       # ground returns inside the corridor only
    ground = np.stack([gx[keep], gy[keep], np.full(keep.sum(), -1.8) + rng.normal(0, 0.005, keep.sum())], 1)
    wx = np.arange(-reach, reach, 0.2)
    left = np.stack([wx, np.full_like(wx, half_width), np.full_like(wx, -1.0)], 1)
    right = np.stack([wx, np.full_like(wx, -half_width), np.full_like(wx, -1.0)], 1)
    parts = [ground, right]
    if gap_at is not None:
        left = left[(wx < gap_at[0]) | (wx > gap_at[1])]
        bx, by = np.meshgrid(np.arange(gap_at[0], gap_at[0] + 2.0, 0.2), np.arange(-1.0, 1.0, 0.2))
        parts.append(np.stack([bx.ravel(), by.ravel(), np.full(bx.size, -1.0)], 1))
    parts.append(left)
    cloud = np.concatenate(parts).astype(np.float32)
    none = np.full(640, 31.0, dtype=np.float32)               # beyond range_max: skipped (LC:111)
    return CostmapScene(none, none.copy(), cloud, (np.float32(-2.26889), np.float32(0.0071), np.float32(30.0)))
