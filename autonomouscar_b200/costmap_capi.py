"""ctypes mirror of include/prius_costmap.h (row N2: the local costmap builder).  No computation."""
import ctypes as C

import numpy as np


class pc_params(C.Structure):
    _fields_ = [("res", C.c_double), ("height_cells", C.c_double), ("width_cells", C.c_double),
                ("z_ground_buffer", C.c_double), ("obstacle_range", C.c_double),
                ("max_inflation_radius", C.c_double), ("global_height_cells", C.c_double),
                ("global_width_cells", C.c_double), ("reserved", C.c_int32 * 8)]


class pc_planar_scan(C.Structure):
    _fields_ = [("ranges", C.POINTER(C.c_float)), ("n", C.c_int32), ("angle_min", C.c_float),
                ("angle_increment", C.c_float), ("range_max", C.c_float), ("roll", C.c_double),
                ("tf_x", C.c_double), ("tf_y", C.c_double), ("z_map", C.c_double),
                ("inv_x", C.c_double), ("inv_y", C.c_double)]


class pc_inputs(C.Structure):
    _fields_ = [("right", pc_planar_scan), ("left", pc_planar_scan), ("cloud_xyz", C.POINTER(C.c_float)),
                ("n_cloud", C.c_int32), ("center_z", C.c_double), ("map_from_base", C.c_double * 12),
                ("base_from_map", C.c_double * 12)]


class pg_params(C.Structure):
    _fields_ = [("res", C.c_double), ("road_width", C.c_double), ("num_lanes", C.c_int32), ("reserved0", C.c_int32),
                ("height_cells", C.c_double), ("width_cells", C.c_double), ("reserved", C.c_int32 * 8)]


def default_global_costmap_params(**overrides):
    """full_sim.launch:68-73 (+ :77 for the resolution)."""
    p = pg_params()
    p.res = 0.25
    p.road_width = 11.1
    p.num_lanes = 2
    p.height_cells = 12000.0
    p.width_cells = 12000.0
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def default_costmap_params(**overrides):
    """full_sim.launch:71-72, 77-82."""
    p = pc_params()
    p.res = 0.25
    p.height_cells = 300.0
    p.width_cells = 300.0
    p.z_ground_buffer = 0.075
    p.obstacle_range = 2.0
    p.max_inflation_radius = 5.0
    p.global_height_cells = 12000.0
    p.global_width_cells = 12000.0
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class CostmapScene:
    """One set of sensor messages + tf numbers; keeps the numpy arrays the pc_inputs point into."""

    def __init__(self, right, left, cloud, scan_meta):
        self.right = np.ascontiguousarray(right, dtype=np.float32)
        self.left = np.ascontiguousarray(left, dtype=np.float32)
        self.cloud = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
        self.scan_meta = scan_meta          # (angle_min, angle_increment, range_max) float32-exact
        self.inputs = pc_inputs()
        fp = C.POINTER(C.c_float)
        for s, arr in ((self.inputs.right, self.right), (self.inputs.left, self.left)):
            s.ranges = arr.ctypes.data_as(fp)
            s.n = len(arr)
            s.angle_min, s.angle_increment, s.range_max = scan_meta
        self.inputs.cloud_xyz = self.cloud.ctypes.data_as(fp)
        self.inputs.n_cloud = len(self.cloud)
        self.tfs = {}

    def set_planar_tf(self, scan, roll, tf_x, tf_y, z_map, inv_x, inv_y):
        scan.roll, scan.tf_x, scan.tf_y, scan.z_map, scan.inv_x, scan.inv_y = roll, tf_x, tf_y, z_map, inv_x, inv_y
