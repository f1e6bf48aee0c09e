"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_ref: the REFERENCE's own local planner
(local_planner.cpp compiled unmodified against the ROS test doubles, see
oracle/ref_planner_harness.cpp and oracle/Makefile).

Only tests/, tests/golden/make_*.py and bench.py's cpu_baseline / --impl reference legs may
import this module.  The .so files are built in the development container (where
/root/reference exists) and travel to the GPU box prebuilt; `available()` says whether they
are there.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from autonomouscar_b200.capi import pp_params, pp_query

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
SHIPPED, AABB = 0, 1
_NAMES = {SHIPPED: "libref_planner_shipped.so", AABB: "libref_planner_aabb.so"}
_REF_SRC = "/root/reference/catkin_ws/src/local_planner/src/local_planner.cpp"


class refp_plan_out(C.Structure):
    _fields_ = [("cap_nodes", C.c_int32), ("cap_path", C.c_int32), ("status", C.c_int32),
                ("n_open", C.c_int32), ("n_closed", C.c_int32), ("n_path", C.c_int32),
                ("n_profile", C.c_int32), ("n_loop_ticks", C.c_int32),
                ("goal_x", C.c_double), ("goal_y", C.c_double),
                ("mp_max_speed", C.c_double), ("mp_max_accel", C.c_double),
                ("open_state", C.POINTER(C.c_double)), ("closed_state", C.POINTER(C.c_double)),
                ("path_xy", C.POINTER(C.c_double)), ("yaw_rates", C.POINTER(C.c_double)),
                ("durations", C.POINTER(C.c_double)), ("speeds", C.POINTER(C.c_double)),
                ("visited", C.POINTER(C.c_uint8))]


def build():
    """(Re)build oracle/_ref when the reference checkout is present; no-op otherwise."""
    if os.path.exists(_REF_SRC):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def available():
    build()
    return all(os.path.exists(os.path.join(_DIR, n)) for n in _NAMES.values())


_libs = {}


def lib(variant):
    if variant not in _libs:
        build()
        L = C.CDLL(os.path.join(_DIR, _NAMES[variant]))
        L.refp_create.restype = C.c_void_p
        L.refp_check_collision.restype = C.c_double
        L.refp_pool_create.restype = C.c_void_p
        L.refp_pool_check_collision.restype = C.c_double
        L.refp_pool_plan_batch.restype = C.c_double
        assert L.refp_variant() == variant
        _libs[variant] = L
    return _libs[variant]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefPlan:
    """What the reference node produced for one LocalNav message."""

    def __init__(self, raw, open_state, closed_state, path_xy, yaw_rates, durations, speeds, visited):
        self.status = raw.status
        self.n_open, self.n_closed = raw.n_open, raw.n_closed
        self.goal = (raw.goal_x, raw.goal_y)
        self.mp_max_speed, self.mp_max_accel = raw.mp_max_speed, raw.mp_max_accel
        self.open_state = open_state[: raw.n_open].copy()
        self.closed_state = closed_state[: raw.n_closed].copy()
        self.path_xy = path_xy[: raw.n_path].copy()
        self.yaw_rates = yaw_rates[: raw.n_profile].copy()
        self.durations = durations[: raw.n_profile].copy()
        self.speeds = speeds[: raw.n_profile].copy()
        self.visited = visited


class RefNode:
    """One Prius::LocalPlanner instance of the reference (variant SHIPPED or AABB)."""

    def __init__(self, params, variant=SHIPPED):
        self.h = None
        self.L = lib(variant)
        self.params = params
        self.variant = variant
        self.h = C.c_void_p(self.L.refp_create(C.byref(params)))
        if not self.h:
            raise RuntimeError("refp_create failed")

    def close(self):
        if self.h:
            self.L.refp_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_base_from_map(self, x, y, yaw):
        self.L.refp_set_base_from_map(C.c_double(x), C.c_double(y), C.c_double(yaw))

    def set_grid_msg(self, data):
        """`data` = nav_msgs/OccupancyGrid.data as received on /local_costmap."""
        data = np.ascontiguousarray(data, dtype=np.int8).ravel()
        rc = self.L.refp_set_grid_msg(self.h, _p(data, C.c_int8), self.params.costmap_width,
                                      self.params.costmap_height)
        assert rc == 0, rc

    def cspace(self):
        out = np.zeros((self.params.costmap_width, self.params.costmap_height), dtype=np.int8)
        self.L.refp_get_cspace(self.h, _p(out, C.c_int8))
        return out

    def footprint(self):
        a, b = C.c_double(), C.c_double()
        self.L.refp_footprint(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def plan(self, query, cap_nodes=1 << 18, cap_path=4096, want_visited=False):
        raw = refp_plan_out()
        raw.cap_nodes, raw.cap_path = cap_nodes, cap_path
        open_state = np.zeros((cap_nodes, 8))
        closed_state = np.zeros((cap_nodes, 8))
        path_xy = np.zeros((cap_path, 2))
        yr, du, sp = np.zeros(cap_path), np.zeros(cap_path), np.zeros(cap_path)
        vis = (np.zeros((self.params.costmap_width, self.params.costmap_height), dtype=np.uint8)
               if want_visited else None)
        raw.open_state, raw.closed_state = _p(open_state, C.c_double), _p(closed_state, C.c_double)
        raw.path_xy = _p(path_xy, C.c_double)
        raw.yaw_rates, raw.durations, raw.speeds = _p(yr, C.c_double), _p(du, C.c_double), _p(sp, C.c_double)
        if vis is not None:
            raw.visited = _p(vis, C.c_uint8)
        rc = self.L.refp_plan(self.h, C.byref(query), C.byref(raw))
        assert rc == 0, rc
        return RefPlan(raw, open_state, closed_state, path_xy, yr, du, sp, vis)

    def possible_velocities(self, query, v):
        out = np.zeros(64)
        n = self.L.refp_possible_velocities(self.h, C.byref(query), C.c_double(v), _p(out, C.c_double), 64)
        return out[:n]

    def possible_yaws(self, heading):
        out = np.zeros(8192)
        n = self.L.refp_possible_yaws(self.h, C.c_double(heading), _p(out, C.c_double), 8192)
        return out[:n]

    def costs(self, query, state6):
        s = np.ascontiguousarray(state6, dtype=np.float64)
        w, h, g = C.c_double(), C.c_double(), C.c_int()
        self.L.refp_costs(self.h, C.byref(query), _p(s, C.c_double), C.byref(w), C.byref(h), C.byref(g))
        return w.value, h.value, bool(g.value)

    def check_collision(self, poses):
        """checkCollision per pose: 0/1 = the reference's answer, 2 = undefined in the reference."""
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 3)
        flags = np.zeros(len(poses), dtype=np.uint8)
        secs = self.L.refp_check_collision(self.h, C.c_int64(len(poses)), _p(poses, C.c_float), _p(flags, C.c_uint8))
        return flags, secs


class RefPool:
    """n_threads reference node instances (one per thread) for the timed CPU arm."""

    def __init__(self, params, grid_msg, n_threads, variant):
        self.h = None
        self.L = lib(variant)
        self.n_threads = n_threads
        g = None if grid_msg is None else np.ascontiguousarray(grid_msg, dtype=np.int8).ravel()
        self.h = C.c_void_p(self.L.refp_pool_create(C.byref(params), None if g is None else _p(g, C.c_int8), n_threads))
        if not self.h:
            raise RuntimeError("refp_pool_create failed")

    def close(self):
        if self.h:
            self.L.refp_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def check_collision(self, poses):
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 3)
        flags = np.zeros(len(poses), dtype=np.uint8)
        secs = self.L.refp_pool_check_collision(self.h, C.c_int64(len(poses)), _p(poses, C.c_float), _p(flags, C.c_uint8))
        return flags, secs

    def plan_batch(self, queries):
        n = len(queries)
        arr = (pp_query * n)(*queries)
        statuses = np.zeros(n, dtype=np.int32)
        n_nodes = np.zeros(n, dtype=np.int32)
        secs = self.L.refp_pool_plan_batch(self.h, n, arr, _p(statuses, C.c_int32), _p(n_nodes, C.c_int32))
        return secs, statuses, n_nodes


def oracle_lists(res):
    """Rebuild m_tree_open / m_tree_closed (vector order) from an oracle / GPU PlanResult:
    open = open slots by id (push_back order, erase keeps order), closed = by closing order."""
    st, cs = res.node_state, res.node_closed_seq
    open_ids = np.nonzero(cs < 0)[0]
    closed_ids = np.nonzero(cs >= 0)[0]
    closed_ids = closed_ids[np.argsort(cs[closed_ids], kind="stable")]
    return st[open_ids], st[closed_ids]


# ---------------------------------------------------------------------------------------------
# row N2: the reference's own local_costmap.cpp (oracle/_ref/libref_costmap.so)
# ---------------------------------------------------------------------------------------------
_costmap_lib = None


def costmap_lib():
    global _costmap_lib
    if _costmap_lib is None:
        build()
        L = C.CDLL(os.path.join(_DIR, "libref_costmap.so"))
        L.refc_create.restype = C.c_void_p
        L.refc_build.restype = C.c_double
        _costmap_lib = L
    return _costmap_lib


def costmap_available():
    build()
    return os.path.exists(os.path.join(_DIR, "libref_costmap.so"))


def costmap_set_tfs(tfs):
    """tfs: (target, source) -> (x, y, z, roll, pitch, yaw); registers them with the tf test double."""
    L = costmap_lib()
    for (target, source), v in tfs.items():
        L.refc_set_tf(target.encode(), source.encode(), *[C.c_double(float(x)) for x in v])


def costmap_derive_inputs(scene):
    """Fill scene.inputs' tf-derived numbers with exactly what the reference node reads from tf."""
    rc = costmap_lib().refc_derive_inputs(C.byref(scene.inputs))
    assert rc == 0, "a tf frame is missing"


class RefCostmap:
    """One Prius::LocalCostmap instance of the reference.  Call costmap_set_tfs first."""

    def __init__(self, params):
        from autonomouscar_b200.costmap_capi import pc_params
        self.h = None
        self.L = costmap_lib()
        self.h = C.c_void_p(self.L.refc_create(C.byref(params)))
        if not self.h:
            raise RuntimeError("refc_create failed")
        self.params = pc_params()
        self.L.refc_get_params(self.h, C.byref(self.params))

    def close(self):
        if self.h:
            self.L.refc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_global(self, grid):
        g = np.ascontiguousarray(grid, dtype=np.int8).ravel()
        self.L.refc_set_global(self.h, _p(g, C.c_int8), C.c_int64(g.size))

    def build(self, scene):
        """right/left LaserScan + centre PointCloud -> the published /local_costmap data, seconds."""
        am, ai, rm = scene.scan_meta
        self.L.refc_set_scan(self.h, 0, _p(scene.right, C.c_float), len(scene.right), C.c_float(am), C.c_float(ai), C.c_float(rm))
        self.L.refc_set_scan(self.h, 1, _p(scene.left, C.c_float), len(scene.left), C.c_float(am), C.c_float(ai), C.c_float(rm))
        n = int(self.params.height_cells * self.params.width_cells)
        out = np.zeros(n, dtype=np.int8)
        secs = self.L.refc_build(self.h, _p(scene.cloud, C.c_float), len(scene.cloud), _p(out, C.c_int8), C.c_int64(n))
        assert secs >= 0, "the node published nothing"
        return out, secs


# ---------------------------------------------------------------------------------------------
# row N3: the reference's own global_costmap.cpp (oracle/_ref/libref_global_costmap.so)
# ---------------------------------------------------------------------------------------------
def global_costmap_available():
    build()
    return os.path.exists(os.path.join(_DIR, "libref_global_costmap.so"))


def write_edges_csv(path, xy, offsets):
    """The edges.csv format read by GlobalCostmap::readEdges (GC:25-50): one polyline per line,
    points "x y 0" separated by commas.  %.9g round-trips every float32 exactly."""
    xy = np.asarray(xy, dtype=np.float32).reshape(-1, 2)
    with open(path, "w") as f:
        for k in range(len(offsets) - 1):
            f.write(",".join("%.9g %.9g 0" % (float(x), float(y)) for x, y in xy[offsets[k]:offsets[k + 1]]) + ",\n")


def global_costmap_build(params, xy, offsets, tmp_dir):
    """Construct Prius::GlobalCostmap on the polylines -> (published int8 grid, seconds)."""
    L = C.CDLL(os.path.join(_DIR, "libref_global_costmap.so"))
    L.refg_build.restype = C.c_double
    path = os.path.join(str(tmp_dir), "edges_for_ref.csv")
    write_edges_csv(path, xy, offsets)
    n = int(params.height_cells * params.width_cells)
    out = np.zeros(n, dtype=np.int8)
    n_out = C.c_int64()
    secs = L.refg_build(C.byref(params), path.encode(), _p(out, C.c_int8), C.c_int64(n), C.byref(n_out))
    assert secs >= 0 and n_out.value == n, (secs, n_out.value, n)
    return out, secs


# ----------------------------------------------------------------------------------------------
# the reference's controller (prius_controller/src/controller.cpp), row N4
# ----------------------------------------------------------------------------------------------
_controller_lib = None


def controller_available():
    build()
    return os.path.exists(os.path.join(_DIR, "libref_controller.so"))


def _controller():
    global _controller_lib
    if _controller_lib is None:
        L = C.CDLL(os.path.join(_DIR, "libref_controller.so"))
        L.refk_create.restype = C.c_void_p
        L.refk_create.argtypes = [C.c_double]
        L.refk_destroy.argtypes = [C.c_void_p]
        L.refk_set_time.argtypes = [C.c_double]
        L.refk_model_state.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.refk_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.refk_step.argtypes = [C.c_void_p]
        L.refk_state.argtypes = [C.c_void_p, C.c_void_p]
        _controller_lib = L
    return _controller_lib


class RefController:
    """The reference's PriusController object behind a manual clock (oracle/ref_controller_harness.cpp)."""

    def __init__(self, t0=0.0):
        self.h = None
        self.L = _controller()
        self.h = self.L.refk_create(t0)

    def close(self):
        if self.h:
            self.L.refk_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def model_state(self, speed, yaw_rate):
        self.L.refk_model_state(self.h, speed, yaw_rate)

    def profile(self, t, yaw_rates, durations, speeds, max_speed, max_accel):
        a, b, c = (np.ascontiguousarray(v, dtype=np.float64) for v in (yaw_rates, durations, speeds))
        self.L.refk_set_time(t)
        self.L.refk_profile(self.h, len(a), a.ctypes.data, b.ctypes.data, c.ctypes.data, max_speed, max_accel)

    def step(self, t):
        """One pass of the 100 Hz loop at clock t; False when the reference would run off the profile."""
        self.L.refk_set_time(t)
        return self.L.refk_step(self.h) == 0

    def state(self):
        """(m_control_it, throttle, brake, steer, shift_gears, messages published)"""
        out = np.zeros(6)
        self.L.refk_state(self.h, out.ctypes.data)
        return tuple(out)
