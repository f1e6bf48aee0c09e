"""TEST INFRASTRUCTURE: prius_msgs/LocalNav (catkin_ws/src/prius_msgs/msg/LocalNav.msg)."""
from _msg import Bag


class LocalNav(Bag):
    def __init__(self):
        super().__init__(x=0.0, y=0.0, speed=0.0, max_speed=0.0, max_accel=0.0)
