"""TEST INFRASTRUCTURE."""
from _msg import Bag


class String(Bag):
    def __init__(self, data=""):
        super().__init__(data=data)
