"""Regenerates tests/golden/ref_planner.npz from oracle/_ref -- the REFERENCE's own
local_planner.cpp, compiled unmodified (see oracle/Makefile, oracle/ref_planner_harness.cpp).
Run in the development container, where /root/reference exists:
    python tests/golden/make_ref_golden.py
The fixture travels to the GPU box; there the oracle and the CUDA path are checked against
these outputs of the reference itself without needing the reference sources.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_COLLISION_REF_AABB, default_params, make_query,  # noqa: E402
                                worlds)

OUT = os.path.dirname(os.path.abspath(__file__))

# (goal_x, goal_y, goal_speed, start_speed, max_expansions, max_nodes)
QUERIES = [
    (20.0, 0.0, 1.5, 0.0, 100000, 20000),    # the survey's probe query
    (20.0, 0.0, 1.5, 1.5, 100000, 20000),
    (20.0, 3.0, 1.5, 1.5, 100000, 20000),
    (12.0, -4.0, 1.5, 0.7, 100000, 20000),
    (25.0, 2.0, 1.5, 0.3, 100000, 20000),
    (8.0, 6.0, 0.0, 1.0, 100000, 20000),     # stop at the goal (velocity 0 allowed, LP:205)
    (20.0, 0.0, 1.5, 0.0, 30, 20000),        # "1 s" elapses after 30 iterations
    (20.0, 5.0, 1.5, 1.5, 100000, 2500),      # ... or when the tree reaches 2500 nodes
]


def msg_grid():
    """nav_msgs/OccupancyGrid.data (message order) of the obstacle world."""
    g = worlds.block_grid(300, 300, frac=0.01, block=4, seed=3)
    return g.reshape(-1)


def params(mode, max_expansions, max_nodes):
    return default_params(collision_mode=mode, max_expansions=max_expansions, max_nodes=max_nodes)


def query(t):
    return make_query(t[0], t[1], goal_speed=t[2], start_speed=t[3])


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def collision_poses():
    rng = np.random.default_rng(2018)
    n = 8192
    return np.stack([rng.uniform(-36, 36, n), rng.uniform(-36, 36, n), rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)


def main():
    from oracle import ref_binding as ref
    assert ref.available(), "oracle/_ref could not be built (is /root/reference present?)"
    blob = {}
    msg = msg_grid()
    for tag, variant, mode in (("shipped", ref.SHIPPED, PP_COLLISION_AS_SHIPPED), ("aabb", ref.AABB, PP_COLLISION_REF_AABB)):
        for k, t in enumerate(QUERIES):
            node = ref.RefNode(params(mode, t[4], t[5]), variant)
            node.set_grid_msg(msg)
            r = node.plan(query(t), want_visited=True)
            pre = f"{tag}{k}_"
            blob[pre + "summary"] = np.array([r.status, r.n_open, r.n_closed, len(r.path_xy), len(r.yaw_rates)], dtype=np.int64)
            # cost of the goal node is never assigned by the reference (LP:289-299): leave it out
            blob[pre + "open_sha"] = sha(r.open_state[:, :7])
            blob[pre + "open_cost_sha"] = sha(r.open_state[:-1, 7] if r.status == 0 else r.open_state[:, 7])
            blob[pre + "closed_sha"] = sha(r.closed_state)
            blob[pre + "closed_state"] = r.closed_state          # closing order = expansion order
            blob[pre + "path_xy"] = r.path_xy
            blob[pre + "yaw_rates"] = r.yaw_rates
            blob[pre + "durations"] = r.durations
            blob[pre + "speeds"] = r.speeds
            blob[pre + "visited"] = np.packbits(r.visited)
            if k == 3:
                blob[pre + "open_state"] = r.open_state
            node.close()
    # grid ingest (LP:425-441) and the dead collision code (LP:446-457, 467-496)
    node = ref.RefNode(params(PP_COLLISION_REF_AABB, 10, 10), ref.AABB)
    node.set_grid_msg(msg)
    blob["cspace_sha"] = sha(node.cspace())
    poses = collision_poses()
    flags, _ = node.check_collision(poses)
    blob["collision_flags"] = flags
    # lattice generators (LP:191-232)
    q = make_query(20.0, 0.0)
    vels = [0.0, 0.225, 0.5, 1.0, 1.275, 1.5, 0.05, 1.45]
    blob["vel_in"] = np.array(vels)
    for k, v in enumerate(vels):
        blob[f"vel{k}"] = node.possible_velocities(q, v)
    headings = np.concatenate([np.arange(-400, 400) * 0.005, np.random.default_rng(7).uniform(-7, 7, 200)])
    blob["yaw_in"] = headings
    blob["yaw_count"] = np.array([len(node.possible_yaws(h)) for h in headings], dtype=np.int32)
    blob["yaw_sha"] = sha(np.concatenate([node.possible_yaws(h) for h in headings]))
    node.close()
    np.savez_compressed(os.path.join(OUT, "ref_planner.npz"), **blob)
    print("wrote ref_planner.npz:", len(blob), "arrays")


if __name__ == "__main__":
    main()
