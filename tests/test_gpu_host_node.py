"""The C++ host-side mirror of the reference node (costmapCallback / gazeboStatesCallback /
localNavCallback -> /mp, /local_path, /goal) compiled against the shared library and run
headless; its outputs are compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from autonomouscar_b200 import PP_COLLISION_REF_AABB, PP_PLAN_FOUND, default_params, make_query

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_node_round_trip(oracle, cuda_device, tmp_path):
    exe = str(tmp_path / "test_host_node")
    pkg = os.path.join(ROOT, "autonomouscar_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "host"),
                           os.path.join(ROOT, "tests", "cpp", "test_host_node.cpp"), "-o", exe,
                           "-L", pkg, "-lprius_planner_b200", f"-Wl,-rpath,{pkg}"])
    out = subprocess.check_output([exe]).decode().split()
    rc0, rc1, errors, n_path, n_mp = (int(v) for v in out[:5])
    gx, gy, ex, ey, last_speed = (float(v) for v in out[5:10])
    n_nodes, n_open, n_closed = (int(v) for v in out[10:13])
    assert rc0 == -1 and errors == 1            # no costmap yet: log, publish nothing (LP:34-39 style)
    assert rc1 == PP_PLAN_FOUND
    assert (gx, gy) == (20.0, 0.0)              # goal transformed map -> base_link (LP:687-690)
    p = default_params(collision_mode=PP_COLLISION_REF_AABB)
    ref = oracle.plan(p, np.zeros((300, 300), np.int8), make_query(20.0, 0.0, start_speed=1.0))
    assert (n_nodes, n_open, n_closed) == (ref.n_nodes, ref.n_open, ref.n_closed)
    assert n_path == ref.n_path and n_mp == ref.n_profile
    assert (ex, ey) == tuple(ref.path_xy[-1]) and last_speed == ref.speeds[-1]
