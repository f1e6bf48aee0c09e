"""Regenerates tests/golden/ref_global_costmap.npz from oracle/_ref/libref_global_costmap.so -- the
REFERENCE's own global_costmap.cpp compiled unmodified (oracle/Makefile).  Run in the development
container: python tests/golden/make_global_costmap_golden.py
Small maps are stored whole (bit-exact targets); the launch-file map (12000 x 12000, built from the
reference's own road polylines, tests/golden/road_edges.npz) is stored as a SHA-256 plus per-row and
per-column road-cell counts."""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from autonomouscar_b200.costmap_capi import default_global_costmap_params  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SMALL = 1200
# (seed, num_lanes, road_width)
SMALL_CASES = [(1, 2, 11.1), (2, 2, 7.3), (3, 3, 11.1)]


def synth_polylines(seed, n_lines=12, ext=149.0):
    """Random-walk polylines on a 300 m map; they run into the border on purpose (clipped strokes)."""
    rng = np.random.default_rng(seed)
    xy, off = [], [0]
    for _ in range(n_lines):
        n = int(rng.integers(1, 30))
        start = rng.uniform(-130, 130, 2)
        pts = np.clip(start + rng.normal(0, 6, (n, 2)).cumsum(0), -ext, ext)
        xy += list(pts)
        off.append(len(xy))
    return np.array(xy, dtype=np.float32), np.array(off, dtype=np.int32)


def small_params(lanes, width):
    return default_global_costmap_params(height_cells=float(SMALL), width_cells=float(SMALL), num_lanes=lanes, road_width=width)


def road_edges():
    z = np.load(os.path.join(OUT, "road_edges.npz"))
    return z["xy"], z["offsets"]


def full_summary(grid, n):
    g = grid.reshape(n, n)
    road = (g >= 0) & (g < 99)
    return (np.frombuffer(hashlib.sha256(grid.tobytes()).digest(), dtype=np.uint8), road.sum(1).astype(np.int32),
            road.sum(0).astype(np.int32), np.array([int(g[road].astype(np.int64).sum())]))


def main():
    from oracle import ref_binding as ref
    assert ref.global_costmap_available()
    blob = {}
    with tempfile.TemporaryDirectory() as d:
        for k, (seed, lanes, width) in enumerate(SMALL_CASES):
            xy, off = synth_polylines(seed)
            grid, _ = ref.global_costmap_build(small_params(lanes, width), xy, off, d)
            blob[f"s{k}_xy"], blob[f"s{k}_off"] = xy, off
            blob[f"s{k}_grid"] = grid
        xy, off = road_edges()
        grid, secs = ref.global_costmap_build(default_global_costmap_params(), xy, off, d)
        blob["full_sha"], blob["full_rows"], blob["full_cols"], blob["full_sum"] = full_summary(grid, 12000)
        blob["full_ref_seconds"] = np.array([secs])
    np.savez_compressed(os.path.join(OUT, "ref_global_costmap.npz"), **blob)
    print("wrote ref_global_costmap.npz; full map: %.2f s in the reference, %d road cells" % (secs, int(blob["full_rows"].sum())))


if __name__ == "__main__":
    main()
