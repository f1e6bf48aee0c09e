"""Builds tests/golden/road_world_4096.npz: BASELINE configs[0] world.

The reference world (prius_sim/worlds/irving.world) has no obstacle models, only 303 road
polylines -- the same polylines as global_costmap/share/edges.csv.
Occupied (= 100) is therefore defined as OFF-ROAD: cell centre farther than half the road
width (11.1 m / 2, full_sim.launch:69) from every polyline segment.  The grid is 4096 x 4096
at 0.1 m in the planner's m_c_space order, expressed in the vehicle frame of the spawn pose
(x = 0, y = 15, yaw = 0: full_sim.launch:59), i.e. cell (2048, 2048) is the car.

Reads /root/reference (read-only) -- run in the build container only:
    python tests/golden/make_road_world.py
The result is committed bit-packed; nothing at test / bench time touches /root/reference.
"""
import os

import numpy as np

EDGES = "/root/reference/catkin_ws/src/global_costmap/share/edges.csv"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "road_world_4096.npz")
N, RES, HALF_WIDTH = 4096, 0.1, 11.1 / 2
SPAWN = (0.0, 15.0)


def main():
    on_road = np.zeros((N, N), dtype=bool)     # [i, j], i <-> -y, j <-> -x (LP:471-472)
    n_seg = 0
    r_cells = int(np.ceil(HALF_WIDTH / RES)) + 1
    for line in open(EDGES):
        pts = []
        for tok in line.strip().split(","):
            v = tok.split()
            if len(v) >= 2:
                pts.append((float(v[0]) - SPAWN[0], float(v[1]) - SPAWN[1]))   # vehicle frame of the spawn pose
        for (x0, y0), (x1, y1) in zip(pts, pts[1:]):
            n_seg += 1
            # cell-coordinate bounding box of the segment, dilated by the half width
            i_lo = int(np.floor(N / 2 - max(y0, y1) / RES)) - r_cells
            i_hi = int(np.ceil(N / 2 - min(y0, y1) / RES)) + r_cells
            j_lo = int(np.floor(N / 2 - max(x0, x1) / RES)) - r_cells
            j_hi = int(np.ceil(N / 2 - min(x0, x1) / RES)) + r_cells
            i_lo, i_hi, j_lo, j_hi = max(i_lo, 0), min(i_hi, N), max(j_lo, 0), min(j_hi, N)
            if i_lo >= i_hi or j_lo >= j_hi:
                continue
            ii, jj = np.mgrid[i_lo:i_hi, j_lo:j_hi]
            cy = (N / 2 - (ii + 0.5)) * RES      # cell centres in metres
            cx = (N / 2 - (jj + 0.5)) * RES
            dx, dy = x1 - x0, y1 - y0
            L2 = dx * dx + dy * dy
            t = np.clip(((cx - x0) * dx + (cy - y0) * dy) / L2, 0.0, 1.0) if L2 > 0 else np.zeros_like(cx)
            d2 = (cx - (x0 + t * dx)) ** 2 + (cy - (y0 + t * dy)) ** 2
            on_road[i_lo:i_hi, j_lo:j_hi] |= d2 <= HALF_WIDTH ** 2
    occupied = ~on_road
    np.savez_compressed(OUT, packed=np.packbits(occupied), n=np.array([N]), res=np.array([RES]),
                        spawn=np.array(SPAWN), half_width=np.array([HALF_WIDTH]))
    print("segments", n_seg, "on-road fraction %.4f" % on_road.mean(), "file", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
