"""TEST INFRASTRUCTURE -- ctypes binding of the CPU oracle (oracle/build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from autonomouscar_b200.capi import pp_params, pp_query, pp_result
from autonomouscar_b200.results import PlanResult

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")
MATH_LIBM, MATH_DET = 0, 1


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".h", ".cpp"))]
    srcs.append(os.path.join(_HERE, "..", "include", "prius_planner.h"))
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_plan_batch_timed.restype = C.c_double
        L.orc_collide_batch.restype = C.c_double
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def cspace_from_msg(data, width, height):
    data = np.ascontiguousarray(data, dtype=np.int8).ravel()
    out = np.zeros(width * height, dtype=np.int8)
    lib().orc_cspace_from_msg(_p(data, C.c_int8), width, height, _p(out, C.c_int8))
    return out.reshape(width, height)


def plan(params, cspace, query, math_mode=MATH_DET, want_visited=False, **caps):
    res = PlanResult(**caps)
    g = None if cspace is None else np.ascontiguousarray(cspace, dtype=np.int8)
    vis = np.zeros(params.costmap_width * params.costmap_height, dtype=np.uint8) if want_visited else None
    rc = lib().orc_plan(C.byref(params), None if g is None else _p(g, C.c_int8), C.byref(query),
                        math_mode, C.byref(res.raw), None if vis is None else _p(vis, C.c_uint8))
    res.rc = rc
    res.visited = None if vis is None else vis.reshape(params.costmap_width, params.costmap_height)
    return res


def plan_batch_timed(params, cspace, queries, math_mode=MATH_DET, n_threads=1):
    n = len(queries)
    arr = (pp_query * n)(*queries)
    g = None if cspace is None else np.ascontiguousarray(cspace, dtype=np.int8)
    statuses = np.zeros(n, dtype=np.int32)
    n_nodes = np.zeros(n, dtype=np.int32)
    digests = np.zeros(n, dtype=np.uint64)
    secs = lib().orc_plan_batch_timed(C.byref(params), None if g is None else _p(g, C.c_int8), n, arr,
                                      math_mode, n_threads, _p(statuses, C.c_int32),
                                      _p(n_nodes, C.c_int32), _p(digests, C.c_uint64))
    return secs, statuses, n_nodes, digests


def possible_velocities(params, query, v):
    out = np.zeros(64)
    n = lib().orc_possible_velocities(C.byref(params), C.byref(query), C.c_double(v), _p(out, C.c_double), 64)
    return out[:n]


def possible_yaws(params, heading):
    out = np.zeros(8192)
    n = lib().orc_possible_yaws(C.byref(params), C.c_double(heading), _p(out, C.c_double), 8192)
    return out[:n]


def collide_batch(params, cspace, poses, n_threads=1):
    poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 3)
    g = np.ascontiguousarray(cspace, dtype=np.int8)
    flags = np.zeros(len(poses), dtype=np.uint8)
    secs = lib().orc_collide_batch(C.byref(params), _p(g, C.c_int8), C.c_int64(len(poses)),
                                   _p(poses, C.c_float), _p(flags, C.c_uint8), n_threads)
    return flags, secs


def collide_edges(params, cspace, edges):
    edges = np.ascontiguousarray(edges, dtype=np.float32).reshape(-1, 5)
    g = np.ascontiguousarray(cspace, dtype=np.int8)
    flags = np.zeros(len(edges), dtype=np.uint8)
    lib().orc_collide_edges(C.byref(params), _p(g, C.c_int8), C.c_int64(len(edges)), _p(edges, C.c_float),
                            _p(flags, C.c_uint8))
    return flags


def footprint_points(params, x, y, yaw):
    out = np.zeros((4096, 2))
    n = lib().orc_footprint_points(C.byref(params), C.c_double(x), C.c_double(y), C.c_double(yaw),
                                   _p(out, C.c_double), 4096)
    return out[:n]


def lattice_dims(params):
    nu, nv, ok = C.c_int(), C.c_int(), C.c_int()
    lib().orc_lattice_dims(C.byref(params), C.byref(nu), C.byref(nv), C.byref(ok))
    return nu.value, nv.value, bool(ok.value)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def sample_poses(seed, stream, first, n, x_lo, x_hi, y_lo, y_hi):
    out = np.zeros((n, 3), dtype=np.float32)
    lib().orc_sample_poses(C.c_uint64(seed), C.c_uint32(stream), C.c_uint64(first), C.c_int64(n),
                           C.c_float(x_lo), C.c_float(x_hi), C.c_float(y_lo), C.c_float(y_hi),
                           _p(out, C.c_float))
    return out


def nearest(node_xy, query_xy):
    node_xy = np.ascontiguousarray(node_xy, dtype=np.float32).reshape(-1, 2)
    query_xy = np.ascontiguousarray(query_xy, dtype=np.float32).reshape(-1, 2)
    idx = np.zeros(len(query_xy), dtype=np.int32)
    d2 = np.zeros(len(query_xy), dtype=np.float32)
    lib().orc_nearest(C.c_int64(len(node_xy)), _p(node_xy, C.c_float), C.c_int64(len(query_xy)),
                      _p(query_xy, C.c_float), _p(idx, C.c_int32), _p(d2, C.c_float))
    return idx, d2


def nearest_grid(node_xy, query_xy, box, cells_per_side=64):
    """S2 through the oracle's bucket grid (BucketGridNN); must equal nearest() on any input."""
    node_xy = np.ascontiguousarray(node_xy, dtype=np.float32).reshape(-1, 2)
    query_xy = np.ascontiguousarray(query_xy, dtype=np.float32).reshape(-1, 2)
    idx = np.zeros(len(query_xy), dtype=np.int32)
    d2 = np.zeros(len(query_xy), dtype=np.float32)
    lib().orc_nearest_grid(C.c_int64(len(node_xy)), _p(node_xy, C.c_float), C.c_int64(len(query_xy)), _p(query_xy, C.c_float),
                           C.c_float(box[0]), C.c_float(box[1]), C.c_float(box[2]), C.c_float(box[3]), int(cells_per_side),
                           _p(idx, C.c_int32), _p(d2, C.c_float))
    return idx, d2


def find_exact(keys, queries):
    keys = np.ascontiguousarray(keys, dtype=np.float64).reshape(-1, 6)
    queries = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 6)
    idx = np.zeros(len(queries), dtype=np.int32)
    lib().orc_find_exact(C.c_int64(len(keys)), _p(keys, C.c_double), C.c_int64(len(queries)),
                         _p(queries, C.c_double), _p(idx, C.c_int32))
    return idx


def det_sincos(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.zeros_like(x), np.zeros_like(x)
    lib().orc_det_sincos(C.c_int64(x.size), _p(x, C.c_double), _p(s, C.c_double), _p(c, C.c_double))
    return s, c


def det_atan2(y, x):
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    o = np.zeros_like(x)
    lib().orc_det_atan2(C.c_int64(x.size), _p(y, C.c_double), _p(x, C.c_double), _p(o, C.c_double))
    return o


def det_wrap_pi(d):
    d = np.ascontiguousarray(d, dtype=np.float64)
    o = np.zeros_like(d)
    lib().orc_det_wrap_pi(C.c_int64(d.size), _p(d, C.c_double), _p(o, C.c_double))
    return o


def rrt_plan(params, cspace, query, **caps):
    res = PlanResult(**caps)
    g = None if cspace is None else np.ascontiguousarray(cspace, dtype=np.int8)
    res.rc = lib().orc_rrt_plan(C.byref(params), None if g is None else _p(g, C.c_int8), C.byref(query),
                                C.byref(res.raw))
    return res


def costmap_build(params, scene_inputs, global_grid=None, math_mode=MATH_DET):
    """One calcOccGrid pass (local_costmap.cpp:71-80) -> (int8 grid in message order, seconds)."""
    lib().orc_costmap_build.restype = C.c_double
    n = int(params.height_cells * params.width_cells)
    out = np.zeros(n, dtype=np.int8)
    g = None if global_grid is None else np.ascontiguousarray(global_grid, dtype=np.int8).ravel()
    secs = lib().orc_costmap_build(C.byref(params), C.byref(scene_inputs), None if g is None else _p(g, C.c_int8),
                                   C.c_int64(0 if g is None else g.size), math_mode, _p(out, C.c_int8))
    return out, secs


def global_costmap_build(params, xy, offsets, math_mode=MATH_DET):
    """GlobalCostmap::setupRoadMap (global_costmap.cpp:52-61) -> (int8 grid in message order, seconds)."""
    lib().orc_global_costmap_build.restype = C.c_double
    xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    n = int(params.height_cells * params.width_cells)
    out = np.zeros(n, dtype=np.int8)
    secs = lib().orc_global_costmap_build(C.byref(params), _p(xy, C.c_float), _p(offsets, C.c_int32), len(offsets) - 1,
                                          math_mode, _p(out, C.c_int8), C.c_int64(n))
    return out, secs
