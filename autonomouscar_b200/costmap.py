"""Python face of include/prius_costmap.h (row N2): the local costmap node's callbacks
(local_costmap.cpp:360-385) on top of the C ABI.  Every computation is a call into
libprius_planner_b200.so; numpy / torch only own memory."""
import ctypes as C

import numpy as np

from . import _lib
from .capi import PP_GRID_OCCUPANCY_MSG, PP_OK
from .costmap_capi import default_costmap_params, default_global_costmap_params


class LocalCostmap:
    """One costmap builder on one GPU (pc_handle).  Method names follow the reference node:
    globalCostmapCallback <- set_global_grid, centerLaserCallback -> pubOccGrid <- build."""

    def __init__(self, params=None, device=0):
        self._L = _lib.load()
        self.params = params if params is not None else default_costmap_params()
        self.device = int(device)
        self._h = None
        h = C.c_void_p()
        self._check(self._L.pc_create(C.byref(self.params), self.device, C.byref(h)))
        self._h = h
        self.n_cells = int(self.params.height_cells * self.params.width_cells)

    def _check(self, rc):
        if rc != PP_OK:
            from .planner import PlannerError
            raise PlannerError(rc, self._L.pp_last_error().decode())

    def close(self):
        if self._h is not None:
            self._L.pc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_global_grid(self, grid):
        """globalCostmapCallback (LC:382-385); pass None / empty for "no global map yet"."""
        if grid is None:
            self._check(self._L.pc_set_global_grid(self._h, None, 0))
            return
        g = np.ascontiguousarray(grid, dtype=np.int8).ravel()
        self._check(self._L.pc_set_global_grid(self._h, g.ctypes.data_as(C.c_void_p), g.size))

    def adopt_global_grid(self, global_map):
        """Take over a device-resident GlobalRoadMap (row N3) without copying it."""
        ptr, n = global_map.release()
        self._check(self._L.pc_adopt_global_grid(self._h, C.c_void_p(ptr), n))

    def build(self, scene, want_host=True):
        """One calcOccGrid pass over `scene` (costmap_capi.CostmapScene).  Returns the int8 grid in
        message order (numpy) or None when only the device grid is wanted."""
        out = np.empty(self.n_cells, dtype=np.int8) if want_host else None
        self._check(self._L.pc_build(self._h, C.byref(scene.inputs),
                                     out.ctypes.data_as(C.c_void_p) if want_host else None))
        return out

    def build_batch(self, scenes, want_host=True, want_masks=True):
        """n independent scenes (one per simulated car) in ONE set of kernel launches.  Returns
        (grids, masks): grids = (n, n_cells) int8 in message order or None, masks = the n region masks of
        front_regions() (uint32) or None.  Identical to n build() calls."""
        from .costmap_capi import pc_inputs
        n = len(scenes)
        arr = (pc_inputs * n)(*[sc.inputs for sc in scenes])       # (copies the structs; they point into the scenes' arrays)
        grids = np.empty((n, self.n_cells), dtype=np.int8) if want_host else None
        masks = np.zeros(n, dtype=np.uint32) if want_masks else None
        self._check(self._L.pc_build_batch(self._h, n, arr, grids.ctypes.data_as(C.c_void_p) if want_host else None,
                                           masks.ctypes.data_as(C.c_void_p) if want_masks else None))
        return grids, masks

    def batch_grids_dev(self):
        """(device pointer, n_scenes) of the most recent batch's grids, scene-major."""
        p, n = C.c_void_p(), C.c_int32()
        self._check(self._L.pc_batch_grids_dev(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def feed_planner_bank(self, planner):
        """Hand the most recent batch's grids to a Planner as a grid bank, without leaving the GPU."""
        ptr, n = self.batch_grids_dev()
        planner.set_grid_bank_dev(ptr, n, layout=PP_GRID_OCCUPANCY_MSG)
        return n

    def grid_dev(self):
        """(device pointer, width, height) of the most recent grid."""
        p, w, h = C.c_void_p(), C.c_int32(), C.c_int32()
        self._check(self._L.pc_grid_dev(self._h, C.byref(p), C.byref(w), C.byref(h)))
        return p.value, w.value, h.value

    def feed_planner(self, planner):
        """Hand the device grid to a Planner without leaving the GPU (/local_costmap -> costmapCallback)."""
        p, w, h = self.grid_dev()
        planner._check(planner._L.pp_set_grid_dev(planner._h, C.c_void_p(p), w, h, self.params.res, PP_GRID_OCCUPANCY_MSG))

    def front_regions(self):
        """Region mask of the most recent grid as the query generator reads it (local_navigator.py:81-257,
        bits PC_REGION_* of prius_costmap.h; decode with rollout.RegionFlags)."""
        m = C.c_uint32()
        self._check(self._L.pc_front_regions(self._h, C.byref(m)))
        return m.value

    def front_regions_of(self, grid_dev_ptr, info_width, n_cells):
        """The same scan over any device-resident OccupancyGrid.data (e.g. a torch int8 tensor's data_ptr())."""
        m = C.c_uint32()
        self._check(self._L.pc_front_regions_dev(self._h, C.c_void_p(grid_dev_ptr), int(info_width), int(n_cells), C.byref(m)))
        return m.value

    def launch_count(self):
        v = C.c_uint64()
        self._check(self._L.pc_launch_count(self._h, C.byref(v)))
        return v.value

    def last_build_ms(self):
        v = C.c_float()
        self._check(self._L.pc_last_build_ms(self._h, C.byref(v)))
        return v.value


class GlobalRoadMap:
    """The static /global_costmap built on the device (pg_build, row N3: GlobalCostmap::setupRoadMap,
    global_costmap.cpp:52-61).  `xy` = float32 (x, y) pairs as read from edges.csv, polyline k =
    points [offsets[k], offsets[k+1])."""

    def __init__(self, xy, offsets, params=None, device=0, want_host=True):
        self._L = _lib.load()
        self.params = params if params is not None else default_global_costmap_params()
        self.device = int(device)
        xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        self.n_cells = int(self.params.height_cells * self.params.width_cells)
        self.host = np.empty(self.n_cells, dtype=np.int8) if want_host else None
        dev = C.c_void_p()
        rc = self._L.pg_build(C.byref(self.params), self.device, xy.ctypes.data_as(C.c_void_p),
                              offsets.ctypes.data_as(C.c_void_p), len(offsets) - 1,
                              self.host.ctypes.data_as(C.c_void_p) if want_host else None, C.byref(dev))
        if rc != PP_OK:
            from .planner import PlannerError
            raise PlannerError(rc, self._L.pp_last_error().decode())
        self._dev = dev.value
        ms, n = C.c_float(), C.c_int32()
        self._L.pg_last_build(C.byref(ms), C.byref(n))
        self.build_ms, self.launches = ms.value, n.value

    def release(self):
        """(device pointer, length); the caller now owns the allocation."""
        ptr, self._dev = self._dev, None
        return ptr, self.n_cells

    def close(self):
        if self._dev:
            self._L.pg_free(self.device, C.c_void_p(self._dev))
            self._dev = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
