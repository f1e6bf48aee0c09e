"""Row N4: closed-loop route roll-outs (autonomouscar_b200/rollout.py).  The loop around the planner
(way-point selection, profile follower, ideal tracker) is host logic; what is checked is that the
SAME loop driven by the reference node, by the oracle and by the CUDA planner drives the same route."""
import numpy as np
import pytest

from autonomouscar_b200 import PP_COLLISION_AS_SHIPPED, PP_PLAN_FOUND, default_params, rollout, worlds


def _oracle_batch(oracle, p, math_mode):
    def run(queries):
        out = []
        for q in queries:
            r = oracle.plan(p, None, q, math_mode=math_mode, want_nodes=False, cap_nodes=8)
            out.append((r.status == PP_PLAN_FOUND, r.yaw_rates.copy(), r.durations.copy(), r.speeds.copy()))
        return out
    return run


def _starts(n):
    """n cars staggered along the first (straight) leg of the route, at different speeds."""
    return [(2.0 * k, 15.0 + 0.05 * (k % 5), 0.02 * ((k % 7) - 3), 0.1 * (k % 12)) for k in range(n)]


def test_route_densification_and_waypoint_rule():
    route = rollout.reference_route()
    assert route[0] == [worlds.ROUTE_X[0], worlds.ROUTE_Y[0]] and route[-1] == [worlds.ROUTE_X[-1], worlds.ROUTE_Y[-1]]
    d = np.hypot(*np.diff(np.array(route), axis=0).T)
    assert d.max() <= 1.0 + 1e-9 and len(route) > 250          # local_navigator.py:27-46: <= 1 m apart
    sel = rollout.WaypointSelector(route)
    last = -1
    for x in np.arange(0.0, 150.0, 7.0):                       # a car driving down the first legs
        wx, wy = sel.select(float(x), 15.0)
        assert sel.path_index >= last                          # the index never goes back (:590)
        last = sel.path_index
        assert np.hypot(wx - x, wy - 15.0) >= 20.0             # :603-609
    sel.select(171.0, 120.0)                                   # next to the end of the route: last way-point, no overrun
    assert sel.path_index == len(route) - 1


def test_profile_follower_cursor():
    f = rollout.ProfileFollower()
    assert f.command(0.0) is None
    f.new_profile(1.0, [0.1, 0.2, 0.3], [125.0, 125.0, 125.0], [1.0, 1.1, 1.2])
    seen = [f.command(1.0 + k / 64.0) for k in range(30)]      # binary-exact times: 8 steps per 125 ms entry
    # controller.cpp:148-153: the cursor moves on when the elapsed time reaches durations[it] / 1000
    assert seen[0] == (0.1, 1.0) and seen[7] == (0.1, 1.0) and seen[8] == (0.2, 1.1) and seen[15] == (0.2, 1.1)
    assert seen[16] == (0.3, 1.2) and seen[23] == (0.3, 1.2)
    assert seen[24] is None and seen[29] is None               # end of the profile (:154-157)


def test_closed_loop_with_reference_node_equals_closed_loop_with_oracle(oracle):
    """The reference's own planner node (oracle/_ref) inside the loop vs the oracle inside the loop."""
    from oracle import ref_binding as ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000)
    node = ref.RefNode(p, ref.SHIPPED)

    def ref_batch(queries):
        out = []
        for q in queries:
            r = node.plan(q, cap_nodes=1 << 15)
            out.append((r.status == 0, r.yaw_rates, r.durations, r.speeds))
        return out
    a = rollout.RouteRollout(_starts(2)).run(ref_batch, 12)
    b = rollout.RouteRollout(_starts(2)).run(_oracle_batch(oracle, p, oracle.MATH_LIBM), 12)
    np.testing.assert_array_equal(a, b)
    assert a[-1, 0, 0] > a[0, 0, 0] and a[-1, 1, 3] > 0        # both cars drive
    node.close()


def test_cars_are_independent(oracle):
    p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000)
    fn = _oracle_batch(oracle, p, oracle.MATH_DET)
    both = rollout.RouteRollout(_starts(3)).run(fn, 6)
    for k in range(3):
        alone = rollout.RouteRollout([_starts(3)[k]]).run(fn, 6)
        np.testing.assert_array_equal(both[:, k], alone[:, 0])


@pytest.mark.gpu
def test_cuda_rollout_equals_oracle_rollout(oracle, cuda_device):
    """64 cars, one pp_plan_batch call per 20 Hz tick: the traces are the oracle's, bit for bit."""
    from autonomouscar_b200 import Planner
    p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000)
    pl = Planner(p, device=0)
    ro = rollout.RouteRollout(_starts(64))
    got = ro.run(rollout.plan_batch_cuda(pl), 10)
    want = rollout.RouteRollout(_starts(64)).run(_oracle_batch(oracle, p, oracle.MATH_DET), 10)
    np.testing.assert_array_equal(got, want)
    assert ro.found == ro.plans == 640
    pl.close()


@pytest.mark.gpu
def test_native_rollout_equals_python_rollout(cuda_device, tmp_path):
    """autonomouscar_b200/host/route_rollout_b200.hpp (C++ on the C ABI) drives the same cars to the same
    states, bit for bit, as the Python loop around the same CUDA planner."""
    import os
    import subprocess
    from autonomouscar_b200 import Planner
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "autonomouscar_b200")
    exe = str(tmp_path / "test_route_rollout")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "include"), "-I", os.path.join(pkg, "host"),
                           os.path.join(root, "tests", "cpp", "test_route_rollout.cpp"), "-o", exe,
                           "-L", pkg, "-lprius_planner_b200", f"-Wl,-rpath,{pkg}"])
    lines = subprocess.check_output([exe, "64", "10", "2.0"]).decode().splitlines()
    secs, plans, found = lines[0].split()
    assert int(plans) == 640 and int(found) == 640
    got = np.array([[float.fromhex(v) for v in ln.split()] for ln in lines[1:]])
    pl = Planner(default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000), device=0)
    want = rollout.RouteRollout(_starts(64)).run(rollout.plan_batch_cuda(pl), 10)[-1]
    np.testing.assert_array_equal(got, want)
    pl.close()
