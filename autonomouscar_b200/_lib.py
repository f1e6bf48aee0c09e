"""Loader of the C-ABI shared library (libprius_planner_b200.so, built in-tree by
`__graft_entry__.build()` / `make -C autonomouscar_b200/csrc`).

There is no fallback of any kind: if the library is missing or does not export a symbol of
include/prius_planner.h, importing the binding raises.
"""
import ctypes as C
import os

from .capi import pp_params, pp_query, pp_result
from .costmap_capi import pc_inputs, pc_params, pg_params

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprius_planner_b200.so")

# every symbol include/prius_planner.h declares
EXPORTS = [
    "pp_default_params", "pp_abi_version", "pp_last_error", "pp_create", "pp_destroy", "pp_get_params",
    "pp_set_collision_mode", "pp_set_grid", "pp_set_grid_dev", "pp_get_cspace", "pp_collide_batch",
    "pp_collide_batch_dev", "pp_collide_batch_bits_dev", "pp_collide_edges_dev", "pp_footprint_points", "pp_sample_poses_dev",
    "pp_philox4x32_10", "pp_nearest_dev", "pp_find_exact_dev", "pp_plan", "pp_plan_batch", "pp_rrt_plan",
    "pp_rrt_plan_batch", "pp_get_visited", "pp_sync", "pp_stream", "pp_launch_count", "pp_last_phase_ms",
    "pp_last_narrow_count", "pp_bench_l2_gather", "pp_set_broad_phase", "pp_group_create", "pp_group_destroy", "pp_group_size",
    "pp_group_set_grid", "pp_group_collide_batch", "pp_group_plan_batch", "pp_set_grid_bank_dev", "pp_plan_batch_banked",
    "pp_pack_flags_dev", "pp_unpack_flags_dev", "pp_plan_record_stride", "pp_plan_batch_dev", "pp_plan_records_unpack",
    "pp_group_records_dev", "pp_collide_batch_q16",
]
# every symbol include/prius_costmap.h declares (row N2)
COSTMAP_EXPORTS = [
    "pc_default_params", "pc_create", "pc_destroy", "pc_set_global_grid", "pc_build", "pc_grid_dev", "pc_stream",
    "pc_launch_count", "pc_last_build_ms", "pc_adopt_global_grid", "pc_front_regions", "pc_front_regions_dev", "pc_build_batch", "pc_batch_grids_dev",
    "pg_default_params", "pg_build", "pg_last_build", "pg_free",
]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    missing = [s for s in EXPORTS + COSTMAP_EXPORTS if not hasattr(L, s)]
    if missing:
        raise LibraryMissing(f"{LIB_PATH} lacks symbols {missing}")
    vp, i32, i64, u64, u32, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_float, C.c_double
    L.pp_last_error.restype = C.c_char_p
    L.pp_default_params.argtypes = [C.POINTER(pp_params)]
    L.pp_create.argtypes = [C.POINTER(pp_params), C.c_int, C.POINTER(vp)]
    L.pp_destroy.argtypes = [vp]
    L.pp_get_params.argtypes = [vp, C.POINTER(pp_params)]
    L.pp_set_collision_mode.argtypes = [vp, C.c_int]
    L.pp_set_grid.argtypes = [vp, vp, i32, i32, f64, C.c_int]
    L.pp_set_grid_dev.argtypes = [vp, vp, i32, i32, f64, C.c_int]
    L.pp_get_cspace.argtypes = [vp, vp, C.c_size_t]
    L.pp_collide_batch.argtypes = [vp, i64, vp, vp]
    L.pp_collide_batch_dev.argtypes = [vp, i64, vp, vp]
    L.pp_collide_batch_bits_dev.argtypes = [vp, i64, vp, vp]
    L.pp_collide_batch_q16.argtypes = [vp, i64, vp, vp, vp]
    L.pp_collide_edges_dev.argtypes = [vp, i64, vp, vp]
    L.pp_pack_flags_dev.argtypes = [vp, i64, vp, vp]
    L.pp_unpack_flags_dev.argtypes = [vp, i64, vp, vp]
    L.pp_footprint_points.argtypes = [vp, f64, f64, f64, vp, i32, C.POINTER(i32)]
    L.pp_sample_poses_dev.argtypes = [vp, u64, u32, u64, i64, f32, f32, f32, f32, vp]
    L.pp_philox4x32_10.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    L.pp_nearest_dev.argtypes = [vp, i64, vp, i64, vp, vp, vp]
    L.pp_find_exact_dev.argtypes = [vp, i64, vp, i64, vp, vp]
    L.pp_plan.argtypes = [vp, C.POINTER(pp_query), C.POINTER(pp_result)]
    L.pp_plan_batch.argtypes = [vp, i32, C.POINTER(pp_query), C.POINTER(pp_result)]
    L.pp_plan_batch_banked.argtypes = [vp, i32, C.POINTER(pp_query), C.POINTER(pp_result)]
    L.pp_plan_record_stride.argtypes = [i32]
    L.pp_plan_record_stride.restype = C.c_size_t
    L.pp_plan_batch_dev.argtypes = [vp, i32, C.POINTER(pp_query), i32, vp]
    L.pp_plan_records_unpack.argtypes = [vp, i32, i32, C.POINTER(pp_result)]
    L.pp_group_records_dev.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.pp_set_grid_bank_dev.argtypes = [vp, i32, vp, i32, i32, f64, C.c_int]
    L.pp_rrt_plan.argtypes = [vp, C.POINTER(pp_query), C.POINTER(pp_result)]
    L.pp_rrt_plan_batch.argtypes = [vp, i32, C.POINTER(pp_query), C.POINTER(pp_result)]
    L.pp_get_visited.argtypes = [vp, vp, C.c_size_t]
    L.pp_sync.argtypes = [vp]
    L.pp_stream.argtypes = [vp, C.POINTER(vp)]
    L.pp_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.pp_last_phase_ms.argtypes = [vp, C.POINTER(f32), i32]
    L.pp_last_narrow_count.argtypes = [vp, C.POINTER(i64)]
    L.pp_bench_l2_gather.argtypes = [vp, C.c_size_t, C.POINTER(f64)]
    L.pp_set_broad_phase.argtypes = [vp, C.c_int]
    L.pp_group_create.argtypes = [C.POINTER(pp_params), C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
    L.pp_group_destroy.argtypes = [vp]
    L.pp_group_size.argtypes = [vp, C.POINTER(C.c_int)]
    L.pp_group_set_grid.argtypes = [vp, vp, i32, i32, f64, C.c_int]
    L.pp_group_collide_batch.argtypes = [vp, i64, vp, vp]
    L.pp_group_plan_batch.argtypes = [vp, i32, C.POINTER(pp_query), C.POINTER(pp_result)]
    L.pc_default_params.argtypes = [C.POINTER(pc_params)]
    L.pc_create.argtypes = [C.POINTER(pc_params), C.c_int, C.POINTER(vp)]
    L.pc_destroy.argtypes = [vp]
    L.pc_set_global_grid.argtypes = [vp, vp, i64]
    L.pc_build.argtypes = [vp, C.POINTER(pc_inputs), vp]
    L.pc_grid_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]
    L.pc_stream.argtypes = [vp, C.POINTER(vp)]
    L.pc_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.pc_last_build_ms.argtypes = [vp, C.POINTER(f32)]
    L.pc_adopt_global_grid.argtypes = [vp, vp, i64]
    L.pc_front_regions.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.pc_front_regions_dev.argtypes = [vp, vp, i32, i64, C.POINTER(C.c_uint32)]
    L.pc_build_batch.argtypes = [vp, i32, vp, vp, vp]
    L.pc_batch_grids_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(i32)]
    L.pg_default_params.argtypes = [C.POINTER(pg_params)]
    L.pg_build.argtypes = [C.POINTER(pg_params), C.c_int, vp, vp, i32, vp, C.POINTER(vp)]
    L.pg_last_build.argtypes = [C.POINTER(f32), C.POINTER(i32)]
    L.pg_free.argtypes = [C.c_int, vp]
    if L.pp_abi_version() != 1:
        raise LibraryMissing("ABI version mismatch")
    _lib = L
    return L
