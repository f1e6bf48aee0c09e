"""Regenerates tests/golden/ref_costmap.npz from oracle/_ref/libref_costmap.so -- the REFERENCE's
own local_costmap.cpp compiled unmodified (oracle/Makefile, oracle/ref_costmap_harness.cpp).
Run in the development container: python tests/golden/make_costmap_golden.py
The fixture stores the sensor messages, the tf-derived numbers the node read and the grid it
published, so the oracle and the CUDA path can be checked against the reference anywhere."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from autonomouscar_b200 import worlds  # noqa: E402
from autonomouscar_b200.costmap_capi import CostmapScene, default_costmap_params  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
GLOBAL_CELLS = 2000
# (scene seed, car pose in the map frame, global-grid band offset or None for "no global map yet")
CASES = [(1, (30.0, -12.0, 0.7), 17), (2, (-100.0, 55.0, -2.1), None), (3, (120.5, -80.25, 3.0), 5)]
SCAN_FIELDS = ("roll", "tf_x", "tf_y", "z_map", "inv_x", "inv_y")


def params():
    return default_costmap_params(global_height_cells=float(GLOBAL_CELLS), global_width_cells=float(GLOBAL_CELLS))


def global_grid(p, car, offset):
    if offset is None:
        return None
    return worlds.road_like_global_grid(GLOBAL_CELLS, offset, zero_at=worlds.costmap_corner_global_cells(p, *car))


def scene_from_fixture(z, k):
    """Rebuild case k's CostmapScene (messages + the tf numbers the reference node read)."""
    sc = CostmapScene(z[f"c{k}_right"], z[f"c{k}_left"], z[f"c{k}_cloud"], tuple(np.float32(v) for v in z[f"c{k}_meta"]))
    for name, scan in (("right", sc.inputs.right), ("left", sc.inputs.left)):
        for f, v in zip(SCAN_FIELDS, z[f"c{k}_{name}_tf"]):
            setattr(scan, f, float(v))
    sc.inputs.center_z = float(z[f"c{k}_center_z"][0])
    for i in range(12):
        sc.inputs.map_from_base[i] = float(z[f"c{k}_map_from_base"][i])
        sc.inputs.base_from_map[i] = float(z[f"c{k}_base_from_map"][i])
    return sc


def main():
    from oracle import ref_binding as ref
    assert ref.costmap_available(), "oracle/_ref/libref_costmap.so could not be built"
    blob = {}
    p = params()
    for k, (seed, car, offset) in enumerate(CASES):
        ref.costmap_set_tfs(worlds.costmap_tfs(*car))
        node = ref.RefCostmap(p)
        sc = worlds.costmap_scene(seed)
        ref.costmap_derive_inputs(sc)
        g = global_grid(p, car, offset)
        if g is not None:
            node.set_global(g)
        grid, _ = node.build(sc)
        blob[f"c{k}_right"], blob[f"c{k}_left"], blob[f"c{k}_cloud"] = sc.right, sc.left, sc.cloud
        blob[f"c{k}_meta"] = np.array(sc.scan_meta, dtype=np.float32)
        for name, scan in (("right", sc.inputs.right), ("left", sc.inputs.left)):
            blob[f"c{k}_{name}_tf"] = np.array([getattr(scan, f) for f in SCAN_FIELDS])
        blob[f"c{k}_center_z"] = np.array([sc.inputs.center_z])
        blob[f"c{k}_map_from_base"] = np.array(list(sc.inputs.map_from_base))
        blob[f"c{k}_base_from_map"] = np.array(list(sc.inputs.base_from_map))
        blob[f"c{k}_grid"] = grid
        node.close()
    np.savez_compressed(os.path.join(OUT, "ref_costmap.npz"), **blob)
    print("wrote ref_costmap.npz", {k: int((blob[f"c{k}_grid"] == 100).sum()) for k in range(len(CASES))})


if __name__ == "__main__":
    main()
