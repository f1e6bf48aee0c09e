"""CPU tests of the boundary: the shared library loads without a GPU, exports every symbol
include/prius_planner.h declares, the ctypes structs match the C layout, and the library
refuses to work (loudly) when there is no CUDA device -- no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "prius_planner.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from autonomouscar_b200 import _lib
    return _lib.load()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from autonomouscar_b200 import _lib
    declared = _declared_functions()
    assert len(declared) >= 30
    assert sorted(_lib.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_every_declared_costmap_symbol_is_exported(lib):
    from autonomouscar_b200 import _lib
    src = open(os.path.join(ROOT, "include", "prius_costmap.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(p[cg]_[a-z0-9_]+)\s*\(", src)))
    assert sorted(_lib.COSTMAP_EXPORTS) == declared and len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), name


def test_costmap_struct_layout_matches_c():
    from autonomouscar_b200.costmap_capi import pc_inputs, pc_params, pc_planar_scan, pg_params
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "prius_costmap.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(pc_params), sizeof(pc_planar_scan), sizeof(pc_inputs),
         offsetof(pc_planar_scan, roll), offsetof(pc_inputs, center_z), offsetof(pc_inputs, base_from_map),
         sizeof(pg_params), offsetof(pg_params, height_cells));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"),
                               os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    want = [C.sizeof(pc_params), C.sizeof(pc_planar_scan), C.sizeof(pc_inputs), pc_planar_scan.roll.offset,
            pc_inputs.center_z.offset, pc_inputs.base_from_map.offset, C.sizeof(pg_params), pg_params.height_cells.offset]
    assert [int(v) for v in out] == want


def test_struct_layout_matches_c():
    from autonomouscar_b200.capi import pp_params, pp_query, pp_result
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "prius_planner.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(pp_params), sizeof(pp_query), sizeof(pp_result),
         offsetof(pp_params, collision_mode), offsetof(pp_params, rrt_goal_bias),
         offsetof(pp_query, seed), offsetof(pp_result, node_state));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"),
                               os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    want = [C.sizeof(pp_params), C.sizeof(pp_query), C.sizeof(pp_result), pp_params.collision_mode.offset,
            pp_params.rrt_goal_bias.offset, pp_query.seed.offset, pp_result.node_state.offset]
    assert [int(v) for v in out] == want


def test_defaults_match_launch_file(lib):
    from autonomouscar_b200.capi import default_params, pp_params
    c = pp_params()
    assert lib.pp_default_params(C.byref(c)) == 0
    py = default_params()
    for name, _ in pp_params._fields_:
        if name == "reserved":
            continue
        assert getattr(c, name) == getattr(py, name), name
    # full_sim.launch:77-79,90-97
    assert (c.costmap_res, c.costmap_width, c.costmap_height) == (0.25, 300, 300)
    assert (c.time_step_ms, c.velocity_res, c.heading_res, c.heading_diff) == (150.0, 0.1, 0.005, 0.065)
    assert (c.goal_pos_tolerance, c.goal_speed_tolerance, c.search_res, c.max_velocity) == (0.1, 0.5, 0.25, 15.0)
    assert (c.car_length, c.car_width) == (4.7, 1.586)


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from autonomouscar_b200 import Planner, PlannerError
    with pytest.raises(PlannerError) as e:
        Planner()
    assert "no CPU fallback" in str(e.value)


def test_argument_errors_do_not_need_a_device(lib):
    from autonomouscar_b200.capi import default_params
    h = C.c_void_p()
    assert lib.pp_create(None, 0, C.byref(h)) == 1                     # PP_ERR_INVALID
    bad = default_params(costmap_width=0)
    assert lib.pp_create(C.byref(bad), 0, C.byref(h)) == 1
    assert b"empty grid" in lib.pp_last_error()
    wide = default_params(costmap_res=0.01)                           # footprint wider than the tile window
    assert lib.pp_create(C.byref(wide), 0, C.byref(h)) == 5           # PP_ERR_UNSUPPORTED
    assert lib.pp_abi_version() == 1


def test_costmap_argument_errors_do_not_need_a_device(lib):
    from autonomouscar_b200.costmap_capi import default_costmap_params, default_global_costmap_params, pc_params, pg_params
    h = C.c_void_p()
    assert lib.pc_create(None, 0, C.byref(h)) == 1                               # PP_ERR_INVALID
    bad = default_costmap_params(res=0.0)
    assert lib.pc_create(C.byref(bad), 0, C.byref(h)) == 1 and b"res" in lib.pp_last_error()
    huge = default_costmap_params(global_height_cells=1e6, global_width_cells=1e6)
    assert lib.pc_create(C.byref(huge), 0, C.byref(h)) == 1
    dev = C.c_void_p()
    g = default_global_costmap_params(num_lanes=0)
    assert lib.pg_build(C.byref(g), 0, None, None, 0, None, C.byref(dev)) == 1
    g = default_global_costmap_params(height_cells=1200.5)
    assert lib.pg_build(C.byref(g), 0, None, None, 0, None, C.byref(dev)) == 1 and b"whole numbers" in lib.pp_last_error()
    # the batched / region-scan / grid-bank entry points refuse null handles before touching CUDA
    m = C.c_uint32()
    assert lib.pc_build_batch(None, 1, None, None, None) == 1
    assert lib.pc_batch_grids_dev(None, C.byref(dev), None) == 1
    assert lib.pc_front_regions(None, C.byref(m)) == 1 and lib.pc_front_regions_dev(None, None, 300, 90000, C.byref(m)) == 1
    assert lib.pp_set_grid_bank_dev(None, 1, None, 300, 300, 0.25, 0) == 1
    assert lib.pp_plan_batch_banked(None, 0, None, None) == 1
    # defaults = launch file (full_sim.launch:68-82)
    c, d = pc_params(), pg_params()
    assert lib.pc_default_params(C.byref(c)) == 0 and lib.pg_default_params(C.byref(d)) == 0
    py = default_costmap_params()
    for name, _ in pc_params._fields_:
        if name != "reserved":
            assert getattr(c, name) == getattr(py, name), name
    pyg = default_global_costmap_params()
    for name, _ in pg_params._fields_:
        if not name.startswith("reserved"):
            assert getattr(d, name) == getattr(pyg, name), name
    assert (c.res, c.height_cells, c.width_cells, c.max_inflation_radius) == (0.25, 300.0, 300.0, 5.0)
    assert (d.road_width, d.num_lanes, d.height_cells) == (11.1, 2, 12000.0)


def test_product_never_touches_the_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "autonomouscar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), (dirpath, f)
                assert not re.search(r"#\s*include\s*[<\"][^>\"]*oracle", text), (dirpath, f)
                assert "liboracle" not in text and "oracle/build" not in text, (dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libprius_planner_b200.so")], capture_output=True, text=True).stdout
    assert "liboracle" not in out


def test_ros_cmake_configures_without_catkin(tmp_path):
    """ros/CMakeLists.txt guards everything ROS-specific behind find_package(catkin QUIET): on a machine without ROS
    (this image) configuring it succeeds and skips the shims (mirrors the executable target of
    the reference's local_planner/CMakeLists.txt:144)."""
    import shutil
    import subprocess
    cmake = shutil.which("cmake")
    if cmake is None:
        import pytest
        pytest.skip("no cmake in this image")
    out = subprocess.run([cmake, "-S", os.path.join(ROOT, "ros"), "-B", str(tmp_path / "b")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    text = open(os.path.join(ROOT, "ros", "CMakeLists.txt")).read()
    assert "find_package(catkin QUIET" in text
    for node in ("local_planner_b200_node", "local_costmap_b200_node", "global_costmap_b200_node"):
        assert node in text and os.path.exists(os.path.join(ROOT, "ros", node + ".cpp"))
    if "catkin not found" not in out.stdout + out.stderr:
        assert "catkin" in out.stdout.lower()          # a ROS machine: the shims are configured instead


def test_round2_struct_layouts_and_argument_errors(lib):
    """The round-2 additions: pp_plan_record_header / pp_pose_quant layouts as the C compiler sees them, the new
    params field, and the argument checks of the new entry points (no device needed: a null handle is refused)."""
    from autonomouscar_b200.capi import pp_params, pp_pose_quant
    from autonomouscar_b200.results import RECORD_HEADER
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "prius_planner.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(pp_plan_record_header), offsetof(pp_plan_record_header, tree_digest),
         offsetof(pp_plan_record_header, cap_path), sizeof(pp_pose_quant), offsetof(pp_params, goal_test_int_abs),
         sizeof(pp_params));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"),
                               os.path.join(d, "t.c")])
        out = [int(v) for v in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert out == [RECORD_HEADER.itemsize, RECORD_HEADER.fields["tree_digest"][1], RECORD_HEADER.fields["cap_path"][1],
                   C.sizeof(pp_pose_quant), pp_params.goal_test_int_abs.offset, C.sizeof(pp_params)]
    assert out[0] == 64
    null = C.c_void_p(None)
    buf = (C.c_uint8 * 64)()
    assert lib.pp_plan_batch_dev(null, 1, None, 16, buf) != 0
    assert lib.pp_pack_flags_dev(null, 8, buf, buf) == 1 and lib.pp_unpack_flags_dev(null, 8, buf, buf) == 1
    assert lib.pp_collide_batch_q16(null, 4, buf, buf, buf) != 0
    assert lib.pp_plan_records_unpack(None, 2, 16, None) == 1 and lib.pp_plan_records_unpack(None, -1, 16, None) == 1
    assert lib.pp_group_records_dev(null, 0, None) == 1
    assert lib.pp_plan_record_stride(-5) == 0
