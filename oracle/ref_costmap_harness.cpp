// TEST INFRASTRUCTURE -- C harness around the REFERENCE's own local costmap node.
//
// #includes /root/reference/catkin_ws/src/local_costmap/src/local_costmap.cpp (REF_LC_CPP) from
// where it lies and compiles it unmodified against the ROS test doubles (oracle/ros_doubles/);
// oracle/Makefile puts the result in oracle/_ref/libref_costmap.so.  The harness delivers the
// messages the node subscribes to (LC:7-10), registers the tf frames it looks up, and returns
// the nav_msgs/OccupancyGrid it publishes on /local_costmap.  No costmap arithmetic lives here.
//
// Undefined behaviour of the reference the caller must steer around: every occ_grid.data[location]
// of LC:131-151, 178-201, 250-254 is unchecked -- keep scan points inside the window, and keep the
// global map 0 where the window's first sample point lands (its local index is -1, LC:263).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdlib.h>
#include <string>
#include <utility>
#include <vector>

#include "../include/prius_costmap.h"
#include "ros_doubles/ros_doubles.h"
#include "ros_doubles/tf/tf.h"

#define private public
#include REF_LC_CPP
#undef private

namespace {
using Prius::LocalCostmap;
struct RefCostmapNode { std::unique_ptr<LocalCostmap> node; };

tf::Transform make_tf(double x, double y, double z, double roll, double pitch, double yaw) {
  tf::Quaternion q;
  q.setRPY(roll, pitch, yaw);
  return tf::Transform(q, tf::Vector3(x, y, z));
}
void mat12(const tf::Transform& t, double* m) {
  for (int r = 0; r < 3; ++r) {
    m[4 * r + 0] = t.getBasis()[r].x(); m[4 * r + 1] = t.getBasis()[r].y(); m[4 * r + 2] = t.getBasis()[r].z();
  }
  m[3] = t.getOrigin().x(); m[7] = t.getOrigin().y(); m[11] = t.getOrigin().z();
}
}  // namespace

extern "C" {

// register tf(target <- source); call for every frame pair before refc_create / refc_build
void refc_set_tf(const char* target, const char* source, double x, double y, double z, double roll, double pitch,
                 double yaw) {
  tf::register_transform(target, source, make_tf(x, y, z, roll, pitch, yaw));
}

void* refc_create(const pc_params* p) {
  auto& prm = ros_doubles::state().params;
  prm.clear();
  prm["local_costmap_res"] = p->res;
  prm["local_costmap_height"] = p->height_cells * p->res;
  prm["local_costmap_width"] = p->width_cells * p->res;
  prm["z_ground_buffer"] = p->z_ground_buffer;
  prm["obstacle_range"] = p->obstacle_range;
  prm["max_inflation_radius"] = p->max_inflation_radius;
  prm["global_costmap_height"] = p->global_height_cells * p->res;
  prm["global_costmap_width"] = p->global_width_cells * p->res;
  auto* r = new RefCostmapNode;
  ros::NodeHandle nh, pnh("~");
  try { r->node.reset(new LocalCostmap(nh, pnh)); } catch (...) { delete r; return nullptr; }
  return r;
}
void refc_destroy(void* h) { delete static_cast<RefCostmapNode*>(h); }

// the cell counts etc. as the node derived them (LC:20-26): res H W Hg Wg
void refc_get_params(void* h, pc_params* out) {
  LocalCostmap& n = *static_cast<RefCostmapNode*>(h)->node;
  out->res = n.m_local_costmap_res;
  out->height_cells = n.m_local_costmap_height;
  out->width_cells = n.m_local_costmap_width;
  out->z_ground_buffer = n.m_z_ground_buffer;
  out->obstacle_range = n.m_obs_range;
  out->max_inflation_radius = n.m_max_inflation_r;
  out->global_height_cells = n.m_global_costmap_height;
  out->global_width_cells = n.m_global_costmap_width;
}

// which: 0 = /prius/front_right_laser/scan (LC:372-375), 1 = front_left (LC:377-380)
void refc_set_scan(void* h, int which, const float* ranges, int n, float angle_min, float angle_increment,
                   float range_max) {
  auto msg = std::make_shared<sensor_msgs::LaserScan>();
  msg->angle_min = angle_min; msg->angle_increment = angle_increment; msg->range_max = range_max;
  msg->ranges.assign(ranges, ranges + n);
  LocalCostmap& node = *static_cast<RefCostmapNode*>(h)->node;
  if (which == 0) node.rightLaserCallback(msg); else node.leftLaserCallback(msg);
}
void refc_set_global(void* h, const int8_t* data, int64_t len) {
  auto msg = std::make_shared<nav_msgs::OccupancyGrid>();
  msg->data.assign(data, data + len);
  static_cast<RefCostmapNode*>(h)->node->globalCostmapCallback(msg);
}

// centerLaserCallback (LC:366-370) -> the published grid.  Returns wall seconds of the callback.
double refc_build(void* h, const float* cloud_xyz, int n_cloud, int8_t* out, int64_t cap) {
  LocalCostmap& node = *static_cast<RefCostmapNode*>(h)->node;
  auto msg = std::make_shared<sensor_msgs::PointCloud>();
  msg->points.resize(n_cloud);
  for (int k = 0; k < n_cloud; ++k) {
    msg->points[k].x = cloud_xyz[3 * k]; msg->points[k].y = cloud_xyz[3 * k + 1]; msg->points[k].z = cloud_xyz[3 * k + 2];
  }
  auto& cap_msgs = ros_doubles::captured<nav_msgs::OccupancyGrid>();
  cap_msgs.clear();
  const auto t0 = std::chrono::steady_clock::now();
  node.centerLaserCallback(msg);
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (cap_msgs.empty()) return -1.0;
  const auto& g = cap_msgs.back();
  const int64_t n = std::min<int64_t>(cap, static_cast<int64_t>(g.data.size()));
  std::memcpy(out, g.data.data(), n);
  cap_msgs.clear();
  return secs;
}

// The numbers mapPlanarScan / mapCenterScan / mapRoadVectors read from tf, produced by the SAME
// test-double calls the node makes (LC:89, 99, 108, 141, 164, 217, 248), so that the oracle and
// the CUDA path can be fed exactly what the reference saw.
int refc_derive_inputs(pc_inputs* in) {
  tf::TransformListener l;
  const char* frames[2] = {"front_right_laser_link", "front_left_laser_link"};
  pc_planar_scan* scans[2] = {&in->right, &in->left};
  try {
    for (int k = 0; k < 2; ++k) {
      tf::StampedTransform t, tz;
      l.lookupTransform(frames[k], "/center_laser_link", ros::Time(0), t);
      l.lookupTransform("/map", frames[k], ros::Time(0), tz);
      double roll, pitch, yaw;
      tf::Matrix3x3(t.getRotation()).getRPY(roll, pitch, yaw);
      scans[k]->roll = roll;
      scans[k]->tf_x = t.getOrigin().getX(); scans[k]->tf_y = t.getOrigin().getY();
      scans[k]->z_map = tz.getOrigin().getZ();
      scans[k]->inv_x = t.inverse().getOrigin().getX(); scans[k]->inv_y = t.inverse().getOrigin().getY();
    }
    tf::StampedTransform c, m;
    l.lookupTransform("/center_laser_link", "/map", ros::Time(0), c);
    in->center_z = c.getOrigin().getZ();
    l.lookupTransform("/map", "/base_link", ros::Time(0), m);
    mat12(m, in->map_from_base);
    mat12(m.inverse(), in->base_from_map);
  } catch (tf::TransformException& e) {
    return 1;
  }
  return 0;
}

}  // extern "C"
