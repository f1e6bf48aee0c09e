"""TEST INFRASTRUCTURE: plain attribute bags standing in for the generated ROS message classes."""


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def pose_stamped(x=0.0, y=0.0, z=0.0, qw=1.0, qx=0.0, qy=0.0, qz=0.0):
    return Bag(header=Bag(), pose=Bag(position=Bag(x=x, y=y, z=z), orientation=Bag(w=qw, x=qx, y=qy, z=qz)))
