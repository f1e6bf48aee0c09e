"""Regenerates tests/golden/ref_navigator.npz by running the REFERENCE's own query generator
(local_navigator.py, imported unmodified from /root/reference against the rospy test doubles of
oracle/rospy_doubles/, see oracle/ref_navigator.py).
Run in the development container: python tests/golden/make_navigator_golden.py
The fixture stores the scripted messages (poses, costmaps) and everything the node published or
held after each loop pass, so the mirror in autonomouscar_b200/rollout.py and the CUDA region scan
can be checked against the reference anywhere."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from autonomouscar_b200 import worlds  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N_PASSES = 150
SEEDS = (11, 12)


def route():
    return [(float(x), float(y)) for x, y in zip(worlds.ROUTE_X, worlds.ROUTE_Y)]


def random_grid(rng, w):
    """A costmap with a few cells == 100 (and near-misses: 99, -1) sprinkled over the columns the
    navigator looks at and their surroundings; some grids hit one region only."""
    g = np.zeros((w, w), dtype=np.int8)
    noise = rng.random((w, w)) < 0.03
    g[noise] = rng.choice(np.array([50, 99, -1], dtype=np.int8), size=int(noise.sum()))
    mode = int(rng.integers(0, 6))
    n = 0 if mode == 0 else int(rng.integers(1, 5))
    for _ in range(n):
        if mode in (1, 2):          # somewhere in the scanned window
            x, y = int(rng.integers(115, 185)), int(rng.integers(170, 300))
        elif mode == 3:             # the F column only
            x, y = int(rng.integers(140, 160)), int(rng.integers(180, 300))
        elif mode == 4:             # the last-cell corners of the R / L regions
            x, y = int(rng.choice([139, 179, 138, 178])), int(rng.choice([209, 239, 269, 299, 208, 298]))
        else:                       # outside every region
            x, y = int(rng.integers(0, 110)), int(rng.integers(0, w))
        if x < w and y < w:
            g[x, y] = 100
    return g


def scenario(seed):
    """(route, passes): a car driven roughly along the route with scripted poses and costmaps."""
    rng = np.random.default_rng(seed)
    rt = route()
    dense = worlds.densified_route()
    passes, s = [], 0.0
    for k in range(N_PASSES):
        s += float(rng.uniform(0.0, 1.5))
        i = min(int(s), len(dense) - 40)
        x = dense[i][0] + float(rng.normal(0, 0.8))
        y = dense[i][1] + float(rng.normal(0, 0.8))
        yaw = float(rng.uniform(-math.pi, math.pi))
        p = {"pose": None, "grid": None, "grid_first": bool(rng.integers(0, 2))}
        if rng.random() > 0.15:
            p["pose"] = (x, y, math.cos(yaw / 2), 0.0, 0.0, math.sin(yaw / 2))
        if rng.random() < 0.3:
            w = int(rng.choice([300, 300, 300, 250, 200, 185, 150]))
            p["grid"] = (w, random_grid(rng, w))
        passes.append(p)
    return rt, passes


def run_reference(rt, passes):
    from oracle import ref_navigator as rn
    script = [dict(p, grid=(p["grid"][0], p["grid"][1].reshape(-1).tolist()) if p["grid"] is not None else None)
              for p in passes]
    return rn.run(rt, script)


def pack(published, states):
    pub = np.full((len(published), 5), np.nan)
    for k, m in enumerate(published):
        if m is not None:
            pub[k] = [m["x"], m["y"], m["speed"], m["max_speed"], m["max_accel"]]
    st = np.array([[s["pathIndex"], s["obstacle"], *s["F"], *s["R"], *s["L"]] for s in states], dtype=np.int32)
    avoid = np.array([s["avoid"] for s in states], dtype=np.float64)
    return pub, st, avoid


def main():
    blob = {}
    for seed in SEEDS:
        rt, passes = scenario(seed)
        pub, st, avoid = pack(*run_reference(rt, passes))
        blob[f"s{seed}_pub"], blob[f"s{seed}_state"], blob[f"s{seed}_avoid"] = pub, st, avoid
        poses = np.full((len(passes), 6), np.nan)
        for k, p in enumerate(passes):
            if p["pose"] is not None:
                poses[k] = p["pose"]
        blob[f"s{seed}_poses"] = poses
        blob[f"s{seed}_grid_first"] = np.array([p["grid_first"] for p in passes])
        blob[f"s{seed}_grid_at"] = np.array([k for k, p in enumerate(passes) if p["grid"] is not None], dtype=np.int32)
        for k, p in enumerate(passes):
            if p["grid"] is not None:
                blob[f"s{seed}_grid{k}"] = p["grid"][1]
        print(f"seed {seed}: {int(np.isfinite(pub[:, 0]).sum())} messages, {len(blob[f's{seed}_grid_at'])} costmaps, "
              f"{int(st[:, 1].sum())} passes with obstacle set")
    np.savez_compressed(os.path.join(OUT, "ref_navigator.npz"), **blob)


def passes_from_fixture(z, seed):
    n = len(z[f"s{seed}_poses"])
    grids = {int(k): z[f"s{seed}_grid{int(k)}"] for k in z[f"s{seed}_grid_at"]}
    out = []
    for k in range(n):
        pose = z[f"s{seed}_poses"][k]
        g = grids.get(k)
        out.append({"pose": None if np.isnan(pose[0]) else tuple(float(v) for v in pose),
                    "grid": None if g is None else (g.shape[0], g), "grid_first": bool(z[f"s{seed}_grid_first"][k])})
    return out


if __name__ == "__main__":
    main()
