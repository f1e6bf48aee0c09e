"""Builds tests/golden/road_edges.npz: the 303 road polylines of the reference's
global_costmap/share/edges.csv (identical to the <road> elements of irving.world) as float32
x, y pairs + polyline offsets -- input DATA for the global road-map rasteriser (row N3), so that
tests and bench never read /root/reference.  Run in the build container only:
    python tests/golden/make_road_edges.py"""
import os

import numpy as np

EDGES = "/root/reference/catkin_ws/src/global_costmap/share/edges.csv"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "road_edges.npz")


def main():
    xy, offsets = [], [0]
    for line in open(EDGES):
        for tok in line.rstrip("\n").split(","):
            c = tok.split(" ")
            if len(c) == 3:                       # GC:40
                xy.append((np.float32(c[0]), np.float32(c[1])))
        offsets.append(len(xy))
    np.savez_compressed(OUT, xy=np.array(xy, dtype=np.float32), offsets=np.array(offsets, dtype=np.int32))
    print(len(offsets) - 1, "polylines,", len(xy), "points,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
