"""TEST INFRASTRUCTURE (see ../rospy/__init__.py)."""
