#!/usr/bin/env python3
"""local_navigator_b200.py -- drop-in replacement of the reference's query-generator node
(catkin_ws/src/local_navigator/src/local_navigator.py) for a ROS box with a B200: same node name
('local_navigation', :724), same subscriptions (base_pose_ground_truth, /path, local_costmap, queue 1,
:537-545), same publisher ('local_nav_waypoints', prius_msgs/LocalNav, queue 1, :548) and the same 20 Hz
loop (:529, :554).

What changes: callbackOccGrid's two nested Python loops over the 90 000-cell message (:101-257, 16 ms per
costmap under CPython 3.12) become one upload of the int8 payload and the CUDA region scan
(pc_front_regions_dev in libprius_planner_b200.so): a few tens of microseconds.  Way-point selection and
the obstacle branch are autonomouscar_b200.rollout.LocalNavigator, which is checked against the reference
module itself (tests/test_navigator_ref_pin.py).

rospy does not exist in this repository's container; the node is run unchanged against rospy test doubles
and its publications are compared with what the reference node published for the same messages
(tests/test_gpu_ros_shim.py::test_navigator_shim_publishes_what_the_reference_published).
"""
import os
import sys
import time

import numpy as np
import rospy
from nav_msgs.msg import OccupancyGrid, Odometry, Path
from prius_msgs.msg import LocalNav

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autonomouscar_b200 import LocalCostmap, rollout  # noqa: E402
from autonomouscar_b200.costmap_capi import default_costmap_params  # noqa: E402

STARTUP_DELAY_S = 35          # the reference waits for the simulator to come up (:531)
LOOP_HZ = rollout.PLAN_HZ     # :529


class RegionScanner:
    """Owns the device-side pieces: a pc_handle (stream + result word) and a reusable int8 staging tensor."""

    def __init__(self, device=0):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.builder = LocalCostmap(default_costmap_params(), device=device)
        self.pinned = self.staged = None

    def mask(self, data, info_width):
        torch = self.torch
        n = len(data)
        if self.pinned is None or self.pinned.numel() < n:
            self.pinned = torch.empty(n, dtype=torch.int8).pin_memory()
            self.staged = torch.empty(n, dtype=torch.int8, device=self.device)
        self.pinned[:n].numpy()[:] = np.asarray(data, dtype=np.int8)
        self.staged[:n].copy_(self.pinned[:n], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.builder.front_regions_of(self.staged.data_ptr(), int(info_width), n)


class LocalNavigatorNode:
    def __init__(self, device=0, startup_delay=STARTUP_DELAY_S):
        self.nav = rollout.LocalNavigator()
        self.scanner = RegionScanner(device)
        time.sleep(startup_delay)
        rospy.Subscriber("base_pose_ground_truth", Odometry, self.on_odometry, queue_size=1)
        rospy.Subscriber("/path", Path, self.on_path, queue_size=1)
        rospy.Subscriber("local_costmap", OccupancyGrid, self.on_costmap, queue_size=1)
        self.pub = rospy.Publisher("local_nav_waypoints", LocalNav, queue_size=1)
        self.msg = LocalNav()

    def on_odometry(self, msg):
        p, q = msg.pose.pose.position, msg.pose.pose.orientation
        self.nav.on_odometry(p.x, p.y, q.w, q.x, q.y, q.z)

    def on_path(self, msg):
        self.nav.on_path([(ps.pose.position.x, ps.pose.position.y) for ps in msg.poses])

    def on_costmap(self, msg):
        self.nav.on_costmap(self.scanner.mask(msg.data, msg.info.width))

    def spin(self):
        rate = rospy.Rate(LOOP_HZ)
        while not rospy.is_shutdown():
            goal = self.nav.spin_once()
            if goal is not None:
                m = self.msg
                m.x, m.y, m.speed, m.max_speed, m.max_accel = goal
                self.pub.publish(m)
            rate.sleep()


def main():
    rospy.init_node("local_navigation")
    try:
        LocalNavigatorNode(device=int(os.environ.get("PRIUS_B200_DEVICE", "0"))).spin()
    except rospy.ROSInterruptException:
        pass


if __name__ == "__main__":
    main()
