"""Row N4 (query generator): autonomouscar_b200/rollout.py's LocalNavigator and the region-scan
restatement oracle/navigator_ref.py against the REFERENCE's own local_navigator.py, imported unmodified
against rospy test doubles (oracle/ref_navigator.py):
  * against the committed outputs of the reference module (tests/golden/ref_navigator.npz) -- everywhere;
  * live, on fresh random drives and costmaps -- where /root/reference is present.
Everything is compared exactly: the published LocalNav doubles, the path index, obstacle / F / R / L flags
and the avoidance way-point after every loop pass."""
import os
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_navigator_golden as mng  # noqa: E402

from autonomouscar_b200 import rollout  # noqa: E402
from oracle import navigator_ref  # noqa: E402


def drive_mirror(route, passes, mask_of=navigator_ref.region_mask):
    """The same script through the mirror; returns arrays shaped like mng.pack()."""
    nav = rollout.LocalNavigator()
    nav.on_path(route)
    pub = np.full((len(passes), 5), np.nan)
    state, avoid = [], []
    for k, p in enumerate(passes):
        events = []
        if p["pose"] is not None:
            events.append(lambda p=p: nav.on_odometry(*p["pose"]))
        if p["grid"] is not None:
            w, g = p["grid"]
            ev = lambda w=w, g=g: nav.on_costmap(mask_of(np.asarray(g).reshape(-1), w))   # noqa: E731
            events.insert(0, ev) if p["grid_first"] else events.append(ev)
        for ev in events:
            ev()
        msg = nav.spin_once()
        if msg is not None:
            pub[k] = msg
        state.append([nav.path_index, nav.obstacle, *nav.F, *nav.R, *nav.L])
        avoid.append(nav.avoid)
    return pub, np.array(state, dtype=np.int32), np.array(avoid, dtype=np.float64)


def assert_same(got, want):
    for g, w, what in zip(got, want, ("published LocalNav", "path index / flags", "avoidance way-point")):
        np.testing.assert_array_equal(g, w, err_msg=what)       # NaN rows (nothing published) compare equal


def test_mirror_reproduces_reference_navigator_fixture():
    z = np.load(os.path.join(GOLD, "ref_navigator.npz"))
    for seed in mng.SEEDS:
        passes = mng.passes_from_fixture(z, seed)
        want = (z[f"s{seed}_pub"], z[f"s{seed}_state"], z[f"s{seed}_avoid"])
        assert_same(drive_mirror(mng.route(), passes), want)
        assert np.isfinite(want[0][:, 0]).sum() > 100 and want[1][:, 1].sum() > 5      # both branches exercised
        assert len({tuple(r) for r in want[1][:, 2:6]}) > 4                             # several F patterns


def test_fixture_inputs_are_the_generators():
    """The committed script is what make_navigator_golden.scenario() builds (no hand edits)."""
    z = np.load(os.path.join(GOLD, "ref_navigator.npz"))
    for seed in mng.SEEDS:
        _, passes = mng.scenario(seed)
        stored = mng.passes_from_fixture(z, seed)
        for a, b in zip(passes, stored):
            assert (a["pose"] is None) == (b["pose"] is None) and (a["grid"] is None) == (b["grid"] is None)
            if a["pose"] is not None:
                assert a["pose"] == b["pose"]
            if a["grid"] is not None:
                np.testing.assert_array_equal(a["grid"][1], b["grid"][1])


def test_mirror_matches_live_reference_navigator():
    from oracle import ref_navigator as rn
    if not rn.available():
        pytest.skip("needs /root/reference (the navigator is imported from there)")
    for seed in (101, 102, 103):
        rt, passes = mng.scenario(seed)
        want = mng.pack(*mng.run_reference(rt, passes))
        assert_same(drive_mirror(rt, passes), want)


def test_reference_navigator_known_answers():
    """Hand-checkable cases of the obstacle branch, run through the reference when it is present."""
    g = np.zeros((300, 300), dtype=np.int8)
    g[150, 285] = 100                                   # F4 only: straight on, 22.5 m ahead, speed 2.0
    m = navigator_ref.region_mask(g.reshape(-1), 300)
    f = rollout.RegionFlags(m)
    assert f.F == [False, False, False, True] and f.obstacle and not any(f.R) and not any(f.L)
    nav = rollout.LocalNavigator()
    nav.on_path(mng.route())
    nav.on_pose(10.0, 20.0, 0.0)
    nav.on_costmap(m)
    assert nav.spin_once() == (32.5, 20.0, 2.0, 1.5, 1.5)
    g[170, 190] = 100                                   # any hit in an L region clears `obstacle` again
    f = rollout.RegionFlags(navigator_ref.region_mask(g.reshape(-1), 300))
    assert f.F[3] and not f.obstacle
    g[:] = 0
    g[145, 185] = 100                                   # F1: stop 5 m ahead
    nav.on_pose(0.0, 0.0, np.pi / 2)
    nav.on_costmap(navigator_ref.region_mask(g.reshape(-1), 300))
    x, y, speed, _, _ = nav.spin_once()
    assert abs(x) < 1e-12 and y == 5.0 and speed == 0.0
    assert nav.spin_once() is None                      # no new pose, nothing published (:556, :707)
    from oracle import ref_navigator as rn
    if rn.available():
        pub, st = rn.run(mng.route(), [{"pose": (0.0, 0.0, *rn.yaw_to_quaternion(np.pi / 2)), "grid_first": False,
                                        "grid": (300, g.reshape(-1).tolist())}])
        assert st[0]["F"] == [True, False, False, False] and st[0]["obstacle"]
        assert pub[0]["speed"] == 0.0 and pub[0]["y"] == 5.0


def test_native_navigator_reproduces_reference_fixture(tmp_path):
    """The C++ mirror (host/route_rollout_b200.hpp, used by the native roll-out) through the same script."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_local_navigator")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "include"),
                           "-I", os.path.join(root, "autonomouscar_b200", "host"),
                           os.path.join(root, "tests", "cpp", "test_local_navigator.cpp"), "-o", exe])
    z = np.load(os.path.join(GOLD, "ref_navigator.npz"))
    for seed in mng.SEEDS:
        passes = mng.passes_from_fixture(z, seed)
        rt = mng.route()
        lines = ["route %d " % len(rt) + " ".join("%r %r" % p for p in rt)]
        for p in passes:
            ev = []
            if p["pose"] is not None:
                ev.append("odom " + " ".join(repr(float(v)) for v in p["pose"]))
            if p["grid"] is not None:
                w, g = p["grid"]
                m = "mask %d" % navigator_ref.region_mask(g.reshape(-1), w)
                ev.insert(0, m) if p["grid_first"] else ev.append(m)
            lines += ev + ["spin"]
        out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
        rows = np.array([[float(v) for v in ln.split()] for ln in out.strip().split("\n")])
        assert len(rows) == len(passes)
        pub = np.where(rows[:, :1] == 1, rows[:, 1:6], np.nan)
        assert_same((pub, rows[:, 6:20].astype(np.int32), rows[:, 20:23]),
                    (z[f"s{seed}_pub"], z[f"s{seed}_state"], z[f"s{seed}_avoid"]))


def test_route_constants_are_what_the_reference_publishes():
    """worlds.ROUTE_X / ROUTE_Y (the way-points every roll-out, bench and golden script starts from) against the
    /path message of the reference's own global_path_planner.py, run against the rospy doubles; then the whole
    chain reference path planner -> reference navigator against the mirror fed from the constants."""
    from autonomouscar_b200 import worlds
    from oracle import ref_navigator as rn
    if not rn.available():
        pytest.skip("needs /root/reference")
    msgs = rn.reference_route(passes=3)
    assert len(msgs) == 3 and msgs[0] == msgs[1] == msgs[2]
    assert msgs[0] == list(zip(worlds.ROUTE_X, worlds.ROUTE_Y)) == [tuple(p) for p in mng.route()]
    _, passes = mng.scenario(55)
    want = mng.pack(*mng.run_reference(msgs[0], passes))          # the reference planner's route into the reference navigator
    assert_same(drive_mirror(mng.route(), passes), want)
