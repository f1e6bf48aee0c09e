"""Row N1: ros/local_planner_b200_node.cpp -- the drop-in replacement of the
reference executable -- is compiled UNCHANGED against the ROS test doubles (the same doubles
the reference's own node is compiled against for oracle/_ref), fed a scripted sequence of
/local_costmap, /gazebo/model_states and /local_nav_waypoints messages plus a map->base_link
tf, and everything it publishes on /goal, /local_path and /mp is compared with what the
REFERENCE NODE publishes for the same messages."""
import os
import subprocess

import numpy as np
import pytest

from autonomouscar_b200 import PP_COLLISION_AS_SHIPPED, PP_COLLISION_REF_AABB, default_params, make_query, worlds

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "test_ros_shim")
    pkg = os.path.join(ROOT, "autonomouscar_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "oracle", "ros_doubles"),
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "host"),
                           os.path.join(ROOT, "tests", "cpp", "test_ros_shim.cpp"), "-o", exe,
                           "-L", pkg, "-lprius_planner_b200", f"-Wl,-rpath,{pkg}"])
    return exe


def _parse(text):
    """-> list of dicts, one per delivered LocalNav."""
    out, cur = [], None
    for line in text.splitlines():
        w = line.split()
        if not w:
            continue
        if w[0] == "delivered":
            cur = {"delivered": int(w[1])}
            out.append(cur)
        elif w[0] == "goal":
            cur["goal"] = [(float.fromhex(w[2 + 3 * k]), float.fromhex(w[3 + 3 * k]), w[4 + 3 * k]) for k in range(int(w[1]))]
        elif w[0] == "path":
            cur["n_path_msgs"] = int(w[1])
            if int(w[1]):
                cur["path_frame"] = w[2]
                v = [float.fromhex(x) for x in w[4:]]
                cur["path"] = np.array(v).reshape(int(w[3]), 2)
        elif w[0] == "mp":
            cur["n_mp_msgs"] = int(w[1])
            if int(w[1]):
                cur["mp_max_speed"], cur["mp_max_accel"] = float.fromhex(w[2]), float.fromhex(w[3])
                cur["mp"] = np.array([float.fromhex(x) for x in w[5:]]).reshape(int(w[4]), 3)
    return out


@pytest.mark.parametrize("mode", [PP_COLLISION_AS_SHIPPED, PP_COLLISION_REF_AABB])
def test_ros_shim_publishes_what_the_reference_node_publishes(cuda_device, tmp_path, mode):
    from oracle import ref_binding as ref
    if not ref.available():
        pytest.skip("oracle/_ref not present")
    exe = _build(tmp_path)
    p = default_params(collision_mode=mode, max_expansions=400, max_nodes=20000)
    cs = worlds.block_grid(300, 300, frac=0.006, block=4, seed=5)
    worlds.clear_disc(cs, p, 0.0, 0.0, 4.0)
    msg = np.ascontiguousarray(cs.reshape(-1)[::-1])
    gpath = str(tmp_path / "grid.bin")
    msg.tofile(gpath)
    # (tf base_link<-map x, y, yaw), twist, LocalNav in the MAP frame
    steps = [((0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (20.0, 0.0, 1.5)),
             ((-10.0, -12.0, 0.0), (0.6, 0.8, 0.0), (30.0, 12.0, 1.5)),
             ((3.0, -2.0, 0.4), (1.2, 0.0, 0.0), (14.0, -3.0, 1.0)),
             ((-50.0, 8.0, -0.9), (0.3, 0.1, 0.0), (40.0, 40.0, 1.5)),
             ((0.0, 0.0, 0.0), (1.5, 0.0, 0.0), (2000.0, 0.0, 1.5))]      # far outside: expansion cap, nothing published
    lines = [f"param local_costmap_res {p.costmap_res!r}", "param local_costmap_height 75", "param local_costmap_width 75",
             f"param collision_mode {mode}", f"param max_expansions {p.max_expansions}", f"param max_nodes {p.max_nodes}",
             f"grid {gpath} 300 300 0.25"]
    for tfm, tw, nav in steps:
        lines += ["tf %r %r %r" % tfm, "twist %r %r %r" % tw, "nav %r %r %r 1.5 1.5" % nav]
    scen = tmp_path / "scenario.txt"
    scen.write_text("\n".join(lines) + "\n")
    got = _parse(subprocess.check_output([exe, str(scen)]).decode())
    assert len(got) == len(steps) and all(g["delivered"] == 1 for g in got)

    node = ref.RefNode(p, ref.AABB if mode == PP_COLLISION_REF_AABB else ref.SHIPPED)
    node.set_grid_msg(msg)
    n_found = 0
    for (tfm, tw, nav), g in zip(steps, got):
        node.set_base_from_map(*tfm)
        q = make_query(nav[0], nav[1], goal_speed=nav[2], start_speed=float(np.sqrt(tw[0] ** 2 + tw[1] ** 2 + tw[2] ** 2)))
        r = node.plan(q)
        # /goal: published before the search, in base_link (LP:633-641)
        assert len(g["goal"]) == 1 and g["goal"][0][2] == "base_link"
        np.testing.assert_allclose(g["goal"][0][:2], r.goal, rtol=0, atol=1e-12)
        if r.status != 0:          # LP:34-39: log, publish nothing
            assert g["n_path_msgs"] == 0 and g["n_mp_msgs"] == 0
            continue
        n_found += 1
        assert g["n_path_msgs"] == 1 and g["n_mp_msgs"] == 1 and g["path_frame"] == "base_link"
        assert g["path"].shape == r.path_xy.shape and g["mp"].shape[0] == len(r.yaw_rates)
        np.testing.assert_allclose(g["path"], r.path_xy, rtol=0, atol=1e-9)
        np.testing.assert_allclose(g["mp"][:, 0], r.yaw_rates, rtol=0, atol=1e-7)
        np.testing.assert_array_equal(g["mp"][:, 1], r.durations)
        np.testing.assert_array_equal(g["mp"][:, 2], r.speeds)
        assert (g["mp_max_speed"], g["mp_max_accel"]) == (r.mp_max_speed, r.mp_max_accel)
    assert n_found >= 2 and n_found < len(steps)
    node.close()


def test_ros_shim_accepts_a_float32_message_resolution(cuda_device, tmp_path):
    """nav_msgs/MapMetaData.resolution is float32: a launch-file 0.1 arrives as 0.100000001490116.  The shim forwards
    it to pp_set_grid, which must not reject the grid over the last bits of a double (the reference never looks at
    the message's resolution, LP:425-435).  0.1 m cells, the reference's collision code enabled: the shim still
    publishes what the reference node publishes."""
    from oracle import ref_binding as ref
    if not ref.available():
        pytest.skip("oracle/_ref not present")
    exe = _build(tmp_path)
    p = default_params(costmap_res=0.1, costmap_width=400, costmap_height=400, collision_mode=PP_COLLISION_REF_AABB,
                       max_expansions=300, max_nodes=20000)
    cs = worlds.block_grid(400, 400, frac=0.004, block=6, seed=15)
    worlds.clear_disc(cs, p, 0.0, 0.0, 5.0)
    msg = np.ascontiguousarray(cs.reshape(-1)[::-1])
    gpath = str(tmp_path / "grid01.bin")
    msg.tofile(gpath)
    steps = [((0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (9.0, 0.5, 1.5)),
             ((-2.0, 1.0, 0.2), (0.8, 0.0, 0.0), (8.0, -1.0, 1.5)),
             ((0.0, 0.0, 0.0), (1.2, 0.0, 0.0), (7.0, 2.0, 1.0))]
    lines = ["param local_costmap_res 0.1", "param local_costmap_height 40", "param local_costmap_width 40",
             f"param collision_mode {PP_COLLISION_REF_AABB}", f"param max_expansions {p.max_expansions}",
             f"param max_nodes {p.max_nodes}", f"grid {gpath} 400 400 {float(np.float32(0.1))!r}"]
    for tfm, tw, nav in steps:
        lines += ["tf %r %r %r" % tfm, "twist %r %r %r" % tw, "nav %r %r %r 1.5 1.5" % nav]
    scen = tmp_path / "scenario01.txt"
    scen.write_text("\n".join(lines) + "\n")
    got = _parse(subprocess.check_output([exe, str(scen)]).decode())
    assert len(got) == len(steps)
    node = ref.RefNode(p, ref.AABB)
    node.set_grid_msg(msg)
    n_found = 0
    for (tfm, tw, nav), g in zip(steps, got):
        node.set_base_from_map(*tfm)
        q = make_query(nav[0], nav[1], goal_speed=nav[2], start_speed=float(np.sqrt(tw[0] ** 2 + tw[1] ** 2 + tw[2] ** 2)))
        r = node.plan(q)
        if r.status != 0:
            assert g["n_path_msgs"] == 0 and g["n_mp_msgs"] == 0
            continue
        n_found += 1
        assert g["n_path_msgs"] == 1 and g["n_mp_msgs"] == 1, "the shim rejected the costmap or the plan"
        np.testing.assert_allclose(g["path"], r.path_xy, rtol=0, atol=1e-9)
        np.testing.assert_array_equal(g["mp"][:, 2], r.speeds)
    assert n_found >= 1
    node.close()


def test_ros_costmap_shim_publishes_the_reference_nodes_grid(cuda_device, tmp_path):
    """Row N2: ros/local_costmap_b200_node.cpp, compiled unchanged against the ROS test doubles, fed the
    same LaserScan / PointCloud / OccupancyGrid messages and tf frames as the reference's
    local_costmap node (oracle/_ref/libref_costmap.so): the /local_costmap it publishes is the
    reference's (cells may differ only where a last-ulp sin/cos/atan2 difference crosses a cell border)."""
    import sys
    from oracle import ref_binding as ref
    if not ref.costmap_available():
        pytest.skip("oracle/_ref not present")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_costmap_golden as mcg
    exe = str(tmp_path / "test_ros_costmap_shim")
    pkg = os.path.join(ROOT, "autonomouscar_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "oracle", "ros_doubles"),
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "host"),
                           os.path.join(ROOT, "tests", "cpp", "test_ros_costmap_shim.cpp"), "-o", exe,
                           "-L", pkg, "-lprius_planner_b200", f"-Wl,-rpath,{pkg}"])
    p = mcg.params()
    car = (42.0, -17.5, 1.1)
    tfs = worlds.costmap_tfs(*car)
    g = mcg.global_grid(p, car, 23)
    gpath = str(tmp_path / "global.bin")
    g.tofile(gpath)
    lines = ["param local_costmap_res 0.25", "param local_costmap_height 75", "param local_costmap_width 75",
             "param z_ground_buffer 0.075", "param obstacle_range 2", "param max_inflation_radius 5",
             f"param global_costmap_height {mcg.GLOBAL_CELLS * 0.25!r}", f"param global_costmap_width {mcg.GLOBAL_CELLS * 0.25!r}"]
    for (t, s), v in tfs.items():
        lines.append(f"tf {t} {s} " + " ".join(repr(float(x)) for x in v))
    scenes = [worlds.costmap_scene(41), worlds.costmap_scene(42)]
    lines.append(f"global {gpath} {g.size}")
    for k, sc in enumerate(scenes):
        for name, arr in (("right", sc.right), ("left", sc.left)):
            f = str(tmp_path / f"{name}{k}.bin")
            arr.tofile(f)
            am, ai, rm = sc.scan_meta
            lines.append(f"scan {name} {f} {len(arr)} {float(am)!r} {float(ai)!r} {float(rm)!r}")
        cf = str(tmp_path / f"cloud{k}.bin")
        sc.cloud.tofile(cf)
        lines.append(f"cloud {cf} {len(sc.cloud)} {tmp_path / f'grid{k}.bin'}")
    scen = tmp_path / "costmap_scenario.txt"
    scen.write_text("\n".join(lines) + "\n")
    out = subprocess.check_output([exe, str(scen)]).decode().splitlines()
    pub = [ln.split() for ln in out if ln.startswith("published")]
    assert len(pub) == 2

    ref.costmap_set_tfs(tfs)
    node = ref.RefCostmap(p)
    node.set_global(g)
    for k, sc in enumerate(scenes):
        w = pub[k]
        assert w[1] == "1" and w[2] == "base_link" and (int(w[3]), int(w[4])) == (300, 300)
        assert float.fromhex(w[5]) == 0.25
        # origin: -W*res/2 + tf(base_link <- center_laser_link).x (LC:54-55)
        assert float.fromhex(w[6]) == -300 * 0.25 / 2 + tfs[("base_link", "center_laser_link")][0]
        got = np.fromfile(str(tmp_path / f"grid{k}.bin"), dtype=np.int8)
        want, _ = node.build(sc)
        assert got.shape == want.shape
        assert (got != want).mean() < 1e-3, int((got != want).sum())
        assert (got == 100).sum() > 1000
    node.close()


def test_ros_global_costmap_shim_publishes_the_reference_nodes_map(cuda_device, tmp_path):
    """Row N3: ros/global_costmap_b200_node.cpp, compiled unchanged against the ROS test doubles, reads the
    same edges.csv-format file as the reference's global_costmap node (oracle/_ref) and publishes the same
    /global_costmap, byte for byte."""
    import sys
    from oracle import ref_binding as ref
    if not ref.global_costmap_available():
        pytest.skip("oracle/_ref not present")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_global_costmap_golden as mgg
    exe = str(tmp_path / "test_ros_global_costmap_shim")
    pkg = os.path.join(ROOT, "autonomouscar_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "oracle", "ros_doubles"),
                           "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_ros_global_costmap_shim.cpp"), "-o", exe,
                           "-L", pkg, "-lprius_planner_b200", f"-Wl,-rpath,{pkg}"])
    xy, off = mgg.synth_polylines(31, n_lines=10)
    p = mgg.small_params(2, 11.1)
    want, _ = ref.global_costmap_build(p, xy, off, tmp_path)          # also leaves edges_for_ref.csv in tmp_path
    out = str(tmp_path / "global.bin")
    line = subprocess.check_output([exe, str(tmp_path / "edges_for_ref.csv"), out, "0.25", "11.1", "2",
                                    repr(mgg.SMALL * 0.25), repr(mgg.SMALL * 0.25)]).decode().split()
    assert line[0] == "published" and line[1] == "1" and line[2] == "map"
    assert (int(line[3]), int(line[4])) == (mgg.SMALL, mgg.SMALL) and float.fromhex(line[5]) == 0.25
    assert float.fromhex(line[6]) == -mgg.SMALL / 8 and float.fromhex(line[7]) == -mgg.SMALL / 8      # GC:69-70
    got = np.fromfile(out, dtype=np.int8)
    np.testing.assert_array_equal(got, want)
    assert ((got >= 0) & (got < 99)).sum() > 100000


def test_navigator_shim_publishes_what_the_reference_published(cuda_device):
    """ros/local_navigator_b200.py, run unchanged against the rospy test doubles, fed the scripted messages of
    tests/golden/ref_navigator.npz: every LocalNav it publishes, and its path index / flags / avoidance
    way-point after every loop pass, equal what the REFERENCE node (local_navigator.py, imported unmodified
    when the fixture was made) published and held.  The costmaps are scanned on the device."""
    import sys
    gold = os.path.join(ROOT, "tests", "golden")
    sys.path.insert(0, gold)
    import make_navigator_golden as mng
    from oracle import ref_navigator as rn
    from test_navigator_ref_pin import assert_same
    shim = os.path.join(ROOT, "ros", "local_navigator_b200.py")
    z = np.load(os.path.join(gold, "ref_navigator.npz"))
    for seed in mng.SEEDS:
        passes = mng.passes_from_fixture(z, seed)
        script = [dict(p, grid=(p["grid"][0], p["grid"][1].reshape(-1).tolist()) if p["grid"] is not None else None)
                  for p in passes]
        got = mng.pack(*rn.run_shim(shim, mng.route(), script))
        assert_same(got, (z[f"s{seed}_pub"], z[f"s{seed}_state"], z[f"s{seed}_avoid"]))
