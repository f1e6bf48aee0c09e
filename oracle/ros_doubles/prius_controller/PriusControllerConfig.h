// TEST INFRASTRUCTURE -- forwards to the ROS test doubles (see ros_doubles.h)
#pragma once
#include "../ros_doubles.h"
