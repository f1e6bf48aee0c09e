"""TEST INFRASTRUCTURE: runs the reference's OWN query generator,
  /root/reference/catkin_ws/src/local_navigator/src/local_navigator.py   (UNMODIFIED, imported where it lies)
against the rospy test doubles of oracle/rospy_doubles/ and returns what it publishes on
/local_nav_waypoints.  Used to pin autonomouscar_b200/rollout.py (row N4) and to generate
tests/golden/ref_navigator.npz (tests/golden/make_navigator_golden.py).  Only tests/ may use this
module; /root/reference does not exist on the GPU box, so GPU tests read the committed fixture.

The node's whole life is its constructor (subscribe, then `while not rospy.is_shutdown()` at
20 Hz, local_navigator.py:546-716).  The doubles deliver the scripted messages from inside
Rate.sleep() and end the loop when the script is exhausted; time.sleep(35) (:531) is skipped.
"""
import contextlib
import importlib.util
import os
import sys
import time

REFERENCE_FILE = "/root/reference/catkin_ws/src/local_navigator/src/local_navigator.py"
_DOUBLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rospy_doubles")
_FAKE = ("rospy", "nav_msgs", "nav_msgs.msg", "std_msgs", "std_msgs.msg", "geometry_msgs", "geometry_msgs.msg",
         "prius_msgs", "prius_msgs.msg", "_msg")


def available():
    return os.path.exists(REFERENCE_FILE)


@contextlib.contextmanager
def _doubles():
    saved = {k: sys.modules.pop(k) for k in _FAKE if k in sys.modules}
    sys.path.insert(0, _DOUBLES)
    real_sleep = time.sleep
    time.sleep = lambda _s: None
    try:
        yield
    finally:
        time.sleep = real_sleep
        sys.path.remove(_DOUBLES)
        for k in _FAKE:
            sys.modules.pop(k, None)
        sys.modules.update(saved)


def yaw_to_quaternion(yaw):
    import math
    return math.cos(yaw / 2), 0.0, 0.0, math.sin(yaw / 2)     # w, x, y, z


def _script(route_xy, passes, Path, Odometry, OccupancyGrid):
    script = [[("/path", Path(route_xy))]]
    for p in passes:
        ev = []
        if p.get("pose") is not None:
            ev.append(("base_pose_ground_truth", Odometry(*p["pose"])))
        if p.get("grid") is not None:
            w, data = p["grid"]
            g = ("local_costmap", OccupancyGrid(width=w, height=w, resolution=0.25, data=data))
            ev.insert(0, g) if p.get("grid_first") else ev.append(g)
        script.append(ev)
    return script


def _drive(rospy, passes, start, snapshot_of):
    """Common part of run() / run_shim(): `start()` constructs the node and blocks in its loop until the
    script is exhausted; `snapshot_of(node)` reads its state after every pass."""
    node_box, states = [], []
    real_sub = rospy.Subscriber.__init__

    def sub_init(self, topic, typ, cb, queue_size=None):       # the callbacks are bound methods of the node
        if not node_box:
            node_box.append(cb.__self__)
        real_sub(self, topic, typ, cb, queue_size)

    real_rate_sleep = rospy.Rate.sleep

    def rate_sleep(self):
        if rospy.state.passes >= 2:      # pass 0 saw nothing, pass 1 saw only /path
            states.append(snapshot_of(node_box[0]))
        real_rate_sleep(self)

    rospy.Subscriber.__init__ = sub_init
    rospy.Rate.sleep = rate_sleep
    start()                              # returns when the script is exhausted
    published = [None] * len(passes)
    for k, _topic, fields in rospy.state.published:
        # loop pass k processed script entry k-1; script entry 0 is /path, entry j+1 is passes[j]
        published[k - 2] = fields
    return published, states


def run(route_xy, passes):
    """Drive the reference node.

    route_xy: the /path way-points [(x, y), ...] (delivered once, before the first pass that sees a pose).
    passes:   one dict per 20 Hz loop pass, delivered just before it:
                {"pose": (x, y, qw, qx, qy, qz) | None,     # base_pose_ground_truth
                 "grid": (info_width, flat int sequence) | None,   # local_costmap, delivered AFTER the pose
                 "grid_first": bool}                        # deliver the grid before the pose instead
    Returns (published, states): published[k] = dict of the LocalNav published in pass k or None;
    states[k] = {"pathIndex", "obstacle", "F": [F1..F4], "R": [...], "L": [...], "avoid": (x, y, speed)}
    read from the node object after pass k.
    """
    with _doubles():
        import rospy
        from nav_msgs.msg import OccupancyGrid, Odometry, Path
        rospy.state.reset()
        spec = importlib.util.spec_from_file_location("reference_local_navigator", REFERENCE_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)          # __name__ != "__main__": defines the class only
        rospy.state.script = _script(route_xy, passes, Path, Odometry, OccupancyGrid)

        def snapshot(n):
            return {"pathIndex": n.pathIndex, "obstacle": bool(n.obstacle),
                    "F": [bool(getattr(n, "regionF%d" % k)) for k in (1, 2, 3, 4)],
                    "R": [bool(getattr(n, "regionR%d" % k)) for k in (1, 2, 3, 4)],
                    "L": [bool(getattr(n, "regionL%d" % k)) for k in (1, 2, 3, 4)],
                    "avoid": (n.obstacle_avoid_wpx, n.obstacle_avoid_wpy, n.obstacle_avoid_spd)}
        return _drive(rospy, passes, mod.LocalNavigator, snapshot)      # the node's life is its constructor


PATH_PLANNER_FILE = "/root/reference/catkin_ws/src/global_navigator/src/global_path_planner.py"


def reference_route(passes=2):
    """The /path messages the reference's global_path_planner.py publishes (its talker() run for `passes` loop
    passes against the doubles): list of [(x, y), ...], one per message."""
    with _doubles():
        import rospy
        rospy.state.reset()
        rospy.state.script = [[] for _ in range(max(passes - 1, 0))]
        spec = importlib.util.spec_from_file_location("reference_global_path_planner", PATH_PLANNER_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.talker()
        return [[(ps.pose.position.x, ps.pose.position.y) for ps in fields["poses"]]
                for _k, topic, fields in rospy.state.published if topic == "path"]


def run_shim(shim_file, route_xy, passes, device=0):
    """The same script through this repository's drop-in node (ros/local_navigator_b200.py), run UNCHANGED
    against the same rospy doubles.  Needs a GPU (the shim scans costmaps on the device)."""
    with _doubles():
        import rospy
        from nav_msgs.msg import OccupancyGrid, Odometry, Path
        rospy.state.reset()
        spec = importlib.util.spec_from_file_location("local_navigator_b200_shim", shim_file)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        rospy.state.script = _script(route_xy, passes, Path, Odometry, OccupancyGrid)

        def snapshot(node):
            n = node.nav
            return {"pathIndex": n.path_index, "obstacle": bool(n.obstacle), "F": list(n.F), "R": list(n.R),
                    "L": list(n.L), "avoid": tuple(n.avoid)}
        return _drive(rospy, passes, lambda: mod.LocalNavigatorNode(device=device).spin(), snapshot)
