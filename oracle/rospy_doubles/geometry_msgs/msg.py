"""TEST INFRASTRUCTURE."""
from _msg import Bag, pose_stamped


class Twist(Bag):
    def __init__(self):
        super().__init__(linear=Bag(x=0.0, y=0.0, z=0.0), angular=Bag(x=0.0, y=0.0, z=0.0))


def PoseStamped(*a, **k):
    return pose_stamped(*a, **k)
