"""TEST INFRASTRUCTURE: nav_msgs/{Odometry, OccupancyGrid, MapMetaData, Path} as attribute bags."""
from _msg import Bag, pose_stamped


class MapMetaData(Bag):
    def __init__(self, width=0, height=0, resolution=0.0):
        super().__init__(width=width, height=height, resolution=resolution, map_load_time=0.0,
                         origin=pose_stamped().pose)


class OccupancyGrid(Bag):
    def __init__(self, width=0, height=0, resolution=0.0, data=()):
        super().__init__(header=Bag(), info=MapMetaData(width, height, resolution), data=data)


class Odometry(Bag):
    def __init__(self, x=0.0, y=0.0, qw=1.0, qx=0.0, qy=0.0, qz=0.0):
        super().__init__(header=Bag(), pose=pose_stamped(x, y, 0.0, qw, qx, qy, qz), twist=Bag())


class Path(Bag):
    def __init__(self, xy=()):
        super().__init__(header=Bag(), poses=[pose_stamped(float(x), float(y)) for x, y in xy])
