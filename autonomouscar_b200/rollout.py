"""Closed-loop route roll-outs without Gazebo.

Host-side mirrors of the two small loops either side of the planner, so that "N independent
queries" become N full drives along the reference's route:

  * the query generator: catkin_ws/src/local_navigator/src/local_navigator.py
      densify_route     :27-46    way-points every `distBetweenWP` = 1 m
      WaypointSelector  :586-609  closest way-point from the current index on, then forward
                                  until it is at least `mindistToNextWP` = 20 m away
      RegionFlags       :81-257   which of the F / R / L regions in front of the car hold a cell == 100
                                  (the scan itself runs on the device grid: pc_front_regions)
      LocalNavigator    :13-78, 259-368, 546-716   the node: three callbacks and one 20 Hz loop pass,
                                  with the obstacle branch (slow down / swerve left / stop way-points)
      query constants   :521-527  speed = max_speed = max_accel = 1.5, 20 Hz
  * the profile follower: catkin_ws/src/prius_controller/src/controller.cpp
      ProfileFollower   :144-158, 178-185   cursor through the /mp profile: entry k lasts
                                  durations[k] / 1000 s of wall time, restarted on every new message
      PriusController   :29-189   the node: that cursor + the steering / pedal PIDs -> prius_msgs/Control

The vehicle itself (Gazebo + the two PIDs of controller.cpp:44-132) is replaced by an ideal tracker:
yaw rate and speed are taken from the profile entry under the cursor and integrated at the
controller's 100 Hz (controller_node.cpp:14).  That stand-in is this repository's, not the
reference's; everything it feeds -- every LocalNav goal, every plan, every profile -- goes through
the same code paths as a live system.  The navigator is a rospy node written for Python 2 and cannot run
here as a ROS node, but it parses as Python 3: tests/test_navigator_ref_pin.py imports it UNMODIFIED against
rospy test doubles (oracle/ref_navigator.py) and checks every published LocalNav, the path index and the
region flags against this mirror (live in this container, committed outputs in tests/golden/ elsewhere).
The controller (cursor and PIDs) is checked the same way against the reference's own controller.cpp compiled
unmodified (oracle/_ref/libref_controller.so, tests/test_controller_ref_pin.py); the planner inside the loop is
pinned (tests/test_rollout.py drives the REFERENCE planner node and the oracle through the same loop).

`plan_batch` is any callable (list of pp_query) -> list of (found, yaw_rates, durations_ms, speeds):
the CUDA planner (plan_batch_cuda), the oracle or the reference node (tests).
"""
import math

import numpy as np

from . import worlds
from .capi import PP_PLAN_FOUND, make_query

PLAN_HZ = 20            # local_navigator.py:529
CONTROL_HZ = 100        # controller_node.cpp:14
DIST_BETWEEN_WP = 1.0   # local_navigator.py:521
MIN_DIST_TO_NEXT_WP = 20.0   # local_navigator.py:522
SPEED = MAX_SPEED = MAX_ACCEL = 1.5   # local_navigator.py:524-527


def densify_route(path, spacing=DIST_BETWEEN_WP):
    """local_navigator.py:27-46."""
    out = [list(path[0])]
    for i in range(len(path) - 1):
        (x0, y0), (x1, y1) = path[i], path[i + 1]
        divide = math.sqrt((x1 - x0) ** 2 + (y1 - y0) ** 2) / spacing
        if divide > 1:
            n = math.ceil(divide)
            dx, dy = (x1 - x0) / n, (y1 - y0) / n
            for k in range(1, int(n + 1)):
                out.append([x0 + k * dx, y0 + k * dy])
        else:
            out.append([x1, y1])
    return out


def reference_route():
    """The 12 way-points global_path_planner.py:101-102 publishes on /path, densified."""
    return densify_route(list(zip(worlds.ROUTE_X, worlds.ROUTE_Y)))


class WaypointSelector:
    """local_navigator.py:586-609 (the no-obstacle branch): one instance per car."""

    def __init__(self, route, min_dist=MIN_DIST_TO_NEXT_WP):
        self.route, self.min_dist, self.path_index = route, min_dist, 0

    def select(self, x, y):
        r, n = self.route, len(self.route)
        closest = 5000.0
        for i in range(self.path_index, n):                      # :590-597
            d = math.sqrt((x - r[i][0]) ** 2 + (y - r[i][1]) ** 2)
            if d < closest:
                self.path_index, closest = i, d
            else:
                break
        d = math.sqrt((x - r[self.path_index][0]) ** 2 + (y - r[self.path_index][1]) ** 2)
        while d < self.min_dist:                                  # :603-609
            if self.path_index >= n - 1:                          # (the reference would run off the list here)
                break
            self.path_index += 1
            if self.path_index == n - 1:
                break
            d = math.sqrt((x - r[self.path_index][0]) ** 2 + (y - r[self.path_index][1]) ** 2)
        return r[self.path_index]


# bits of the region mask (include/prius_costmap.h, PC_REGION_*)
def _region_bit(base, k):
    return 1 << (base + k - 1)


class RegionFlags:
    """Decoded pc_front_regions mask: F[k-1], R[k-1], L[k-1] for k = 1..4, `obstacle`, and which of the
    R / L flags the scan rewrote (a grid narrower than 300 cells leaves some untouched)."""

    def __init__(self, mask):
        self.mask = mask
        self.F = [bool(mask & _region_bit(0, k)) for k in (1, 2, 3, 4)]
        self.R = [bool(mask & _region_bit(4, k)) for k in (1, 2, 3, 4)]
        self.L = [bool(mask & _region_bit(8, k)) for k in (1, 2, 3, 4)]
        self.obstacle = bool(mask & (1 << 12))
        self.R_scanned = [bool(mask & _region_bit(16, k)) for k in (1, 2, 3, 4)]
        self.L_scanned = [bool(mask & _region_bit(20, k)) for k in (1, 2, 3, 4)]


def yaw_from_quaternion(qw, qx, qy, qz):
    """local_navigator.py:70-75."""
    return math.atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz))


class LocalNavigator:
    """The query generator node (local_navigator.py) as its three callbacks and one loop pass.

    on_path / on_pose / on_costmap are callbackPath / callbackCarPos / callbackOccGrid; the last takes
    the region mask of the costmap (LocalCostmap.front_regions(), computed on the device) instead of the
    OccupancyGrid message.  spin_once() is one pass of the 20 Hz loop (:551-716) and returns the
    LocalNav it publishes as (x, y, speed, max_speed, max_accel), or None."""

    def __init__(self, spacing=DIST_BETWEEN_WP, min_dist=MIN_DIST_TO_NEXT_WP):
        self.spacing, self.min_dist = spacing, min_dist
        self.selector = None
        self.pose = [0.0, 0.0, 0.0, 0.0]                 # x, y, 0, theta (:76)
        self.got_pose = False
        self.obstacle = False
        self.F, self.R, self.L = [False] * 4, [False] * 4, [False] * 4
        self.avoid = (0.0, 0.0, 0.0)                     # obstacle_avoid_wpx, _wpy, _spd (:498-500)
        self.intermediate_wp = False

    @property
    def path_index(self):
        return self.selector.path_index if self.selector else 0

    def on_path(self, xy):                               # :13-62, acts on the first message only
        if self.selector is None:
            self.selector = WaypointSelector(densify_route([list(p) for p in xy], self.spacing), self.min_dist)

    def on_pose(self, x, y, yaw):                        # :66-78
        self.pose = [x, y, 0, yaw]
        self.got_pose = True

    def on_odometry(self, x, y, qw, qx, qy, qz):
        self.on_pose(x, y, yaw_from_quaternion(qw, qx, qy, qz))

    def on_costmap(self, mask):                          # :81-368
        f = mask if isinstance(mask, RegionFlags) else RegionFlags(mask)
        self.obstacle, self.F = f.obstacle, list(f.F)
        self.R = [v if seen else old for v, seen, old in zip(f.R, f.R_scanned, self.R)]
        self.L = [v if seen else old for v, seen, old in zip(f.L, f.L_scanned, self.L)]
        x, y, heading = self.pose[0], self.pose[1], self.pose[3]
        c, s = math.cos(heading), math.sin(heading)
        cl, sl = math.cos(heading + 1.5708), math.sin(heading + 1.5708)
        ax, ay, spd = self.avoid
        if self.F[3]:                                    # :278-290  far: keep straight, slow down
            ax, ay, spd = x + c * 22.5, y + s * 22.5, 2.0
        if self.F[2]:                                    # :292-305  nearer: pass on the left
            ax, ay, spd = x + c * 15 + cl * 6, y + s * 15 + sl * 6, 1.5
        if self.F[1]:                                    # :307-320
            ax, ay, spd = x + c * 7.5 + cl * 7, y + s * 7.5 + sl * 7, 1.0
        if self.F[0]:                                    # :322-335  directly ahead: stop
            ax, ay, spd = x + c * 5, y + s * 5, 0.0
        self.avoid = (ax, ay, spd)

    def spin_once(self):                                 # :551-716
        if not self.got_pose:
            return None
        self.got_pose = False
        if self.selector is None:
            return None
        if self.obstacle:                                # :561-583
            self.intermediate_wp = True
            return (self.avoid[0], self.avoid[1], self.avoid[2], MAX_SPEED, MAX_ACCEL)
        wx, wy = self.selector.select(self.pose[0], self.pose[1])     # :586-624
        self.intermediate_wp = False
        return (wx, wy, SPEED, MAX_SPEED, MAX_ACCEL)


class ProfileFollower:
    """controller.cpp:144-158 / 178-185: which profile entry is being tracked at time t."""

    def __init__(self):
        self.yaw_rates, self.durations, self.speeds = (), (), ()
        self.it, self.start = 0, 0.0

    def new_profile(self, t, yaw_rates, durations_ms, speeds):     # motionPlanningCallback
        self.yaw_rates, self.durations, self.speeds = yaw_rates, durations_ms, speeds
        self.it, self.start = 0, t

    def command(self, t):
        """(yaw_rate, speed) at time t, or None when there is no entry left ("good luck", :154-157)."""
        if self.it < len(self.durations) and (t - self.start) >= self.durations[self.it] / 1000:   # :148-153
            self.start = t
            self.it += 1
        if self.it >= len(self.yaw_rates):
            return None
        return self.yaw_rates[self.it], self.speeds[self.it]


def _ieee_div(a, b):
    """a / b as C++ doubles do it (the reference divides by dt = 0 when two passes share a clock value)."""
    if b != 0.0:
        return a / b
    if a == 0.0 or a != a:
        return math.nan
    return math.copysign(math.inf, a) * math.copysign(1.0, b)


class PriusController:
    """The whole controller node (controller.cpp): the /mp cursor plus the steering and pedal PIDs, producing
    prius_msgs/Control.  The roll-outs do not need it (they track the profile ideally); it is here so that the
    cursor they share with it is checked against the reference's own object (tests/test_controller_ref_pin.py,
    oracle/_ref/libref_controller.so) and for anyone putting a vehicle model behind the loop.

    Gains are the dynamic_reconfigure defaults (cfg/prius_controller.cfg).  m_speed is uninitialised in the
    reference (controller.hpp:55); it starts at 0 here and in the harness.  step() returns False where the
    reference would index past the end of the profile (:148, :47) -- "good luck" (:154-157)."""
    NO_COMMAND, NEUTRAL, FORWARD, REVERSE = 0, 1, 2, 3

    def __init__(self, t0=0.0, kp_s=7.5, ki_s=0.05, kd_s=0.15, kp_p=0.25, ki_p=0.075, kd_p=0.055):
        self.kp_s, self.ki_s, self.kd_s, self.kp_p, self.ki_p, self.kd_p = kp_s, ki_s, kd_s, kp_p, ki_p, kd_p
        self.cursor = ProfileFollower()
        self.cursor.start = t0
        self.prev_t_s = self.prev_t_p = t0
        self.prev_err_s = self.int_s = self.prev_err_p = self.int_p = 0.0
        self.m_speed = 0.0
        self.state_speed = self.state_yaw_rate = 0.0          # m_prius_state.twist.linear.x / angular.z
        self.max_speed = self.max_accel = 0.0
        self.throttle = self.brake = self.steer = 0.0
        self.shift_gears = self.NO_COMMAND
        self.received = False
        self.published = 0

    @property
    def control_it(self):
        return self.cursor.it

    def on_model_state(self, speed, yaw_rate):                 # :160-174, 185-189
        self.state_speed, self.state_yaw_rate = speed, yaw_rate

    def on_profile(self, t, yaw_rates, durations_ms, speeds, max_speed, max_accel):      # :176-183
        self.received = True
        self.cursor.new_profile(t, yaw_rates, durations_ms, speeds)
        self.max_speed, self.max_accel = max_speed, max_accel
        return self.step(t)

    def step(self, t):                                         # calculateAndPublishControls :29-42
        cmd = self.cursor.command(t)                           # getCurrentCommandIterator :144-158
        if cmd is None:
            return False
        yaw_rate_cmd, speed_cmd = cmd
        # calculateSteering :44-63
        dt = t - self.prev_t_s
        err = yaw_rate_cmd - self.state_yaw_rate
        self.int_s += err * dt
        der = _ieee_div(err - self.prev_err_s, dt)
        steer = self.kp_s * err + self.ki_s * self.int_s - self.kd_s * der
        if steer > 1:
            steer = 1.0
        if steer < -1:
            steer = -1.0
        self.steer, self.prev_t_s, self.prev_err_s = steer, t, err
        # calculateGear :65-76 (m_speed still holds the previous pass's profile speed)
        if self.m_speed > 0:
            self.shift_gears = self.FORWARD
        if self.m_speed < 0:
            self.shift_gears = self.REVERSE
        # calculatePedals :78-132
        dt = t - self.prev_t_p
        self.m_speed = speed_cmd
        speed = self.state_speed
        if speed > self.max_speed:
            speed = self.max_speed
        if self.m_speed < speed:                               # (the `>` branch is overwritten by the else, :90-103)
            set_point = speed - self.max_accel * dt
        else:
            set_point = self.m_speed
        err = set_point - speed
        self.int_p += err * dt
        der = _ieee_div(err - self.prev_err_p, dt)
        u = self.kp_p * err + self.ki_p * self.int_p - self.kd_p * der
        if u > 1:
            u = 1.0
        if u < -1:
            u = -1.0
        if u >= 0:
            self.throttle, self.brake = u, 0.0
        if u < 0:
            self.brake, self.throttle = -u, 0.0
        self.prev_t_p, self.prev_err_p = t, err
        self.published += 1                                    # publishControls :134-137
        return True


class RouteRollout:
    """S independent cars on the reference route, stepped in lock-step at the 20 Hz plan rate."""

    def __init__(self, starts, route=None):
        """starts: iterable of (x, y, yaw, speed) in the map frame."""
        self.route = route if route is not None else reference_route()
        st = np.array(list(starts), dtype=np.float64).reshape(-1, 4)
        self.x, self.y, self.yaw, self.v = st[:, 0].copy(), st[:, 1].copy(), st[:, 2].copy(), st[:, 3].copy()
        self.n = len(st)
        self.navigators = [LocalNavigator() for _ in range(self.n)]
        for nav in self.navigators:                     # the route is already dense: share it
            nav.selector = WaypointSelector(self.route)
        self.followers = [ProfileFollower() for _ in range(self.n)]
        self.t = 0.0
        self.plans = self.found = 0
        self.trace = []

    def queries(self, masks=None):
        """One LocalNav per car, goal converted to its base_link frame (LP:670-691).  masks[k] (optional) is the
        region mask of car k's current costmap (LocalCostmap.front_regions()): with it the navigator's obstacle
        branch may replace the route way-point by a slow-down / swerve / stop way-point."""
        qs = []
        for k, nav in enumerate(self.navigators):
            nav.on_pose(float(self.x[k]), float(self.y[k]), float(self.yaw[k]))
            if masks is not None:
                nav.on_costmap(masks[k])
            wx, wy, speed, max_speed, max_accel = nav.spin_once()
            c, s = math.cos(self.yaw[k]), math.sin(self.yaw[k])
            dx, dy = wx - self.x[k], wy - self.y[k]
            qs.append(make_query(c * dx + s * dy, -s * dx + c * dy, goal_speed=speed, max_speed=max_speed,
                                 max_accel=max_accel, start_speed=float(self.v[k]), query_id=k, grid_index=k))
        return qs

    def tick(self, plan_batch, masks=None):
        results = plan_batch(self.queries(masks))
        self.plans += self.n
        for k, (found, yaw_rates, durations, speeds) in enumerate(results):
            if found:                                   # a failed plan publishes nothing: the old profile stays
                self.found += 1
                self.followers[k].new_profile(self.t, yaw_rates, durations, speeds)
        dt = 1.0 / CONTROL_HZ
        for _ in range(CONTROL_HZ // PLAN_HZ):          # ideal tracking of the profile at the controller rate
            for k in range(self.n):
                cmd = self.followers[k].command(self.t)
                if cmd is None:
                    continue
                self.yaw[k] += cmd[0] * dt
                self.v[k] = cmd[1]
                self.x[k] += self.v[k] * dt * math.cos(self.yaw[k])
                self.y[k] += self.v[k] * dt * math.sin(self.yaw[k])
            self.t += dt
        self.trace.append(np.stack([self.x, self.y, self.yaw, self.v], 1).copy())

    def run(self, plan_batch, ticks):
        for _ in range(ticks):
            self.tick(plan_batch)
        return np.array(self.trace)


def plan_batch_cuda(planner, cap_path=256, banked=False):
    """plan_batch callable on top of pp_plan_batch (one C-ABI call per tick for all cars).  banked=True: car k
    plans on grid k of the planner's grid bank (pp_plan_batch_banked; fill it with LocalCostmap.feed_planner_bank)."""
    from .capi import pp_query
    from .results import PlanBatch
    cache = {}

    def run(queries):
        n = len(queries)
        if n not in cache:
            cache[n] = PlanBatch(n, cap_path=cap_path)
        buf = cache[n]
        planner.plan_batch_into((pp_query * n)(*queries), buf, banked=banked)
        st, npf = buf.field("status"), buf.field("n_profile")
        return [(st[k] == PP_PLAN_FOUND, buf.yaw_rates[k, :npf[k]].copy(), buf.durations[k, :npf[k]].copy(),
                 buf.speeds[k, :npf[k]].copy()) for k in range(n)]
    return run
