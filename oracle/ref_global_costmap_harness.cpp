// TEST INFRASTRUCTURE -- C harness around the REFERENCE's own global costmap node.
//
// #includes /root/reference/catkin_ws/src/global_costmap/src/global_costmap.cpp (REF_GC_CPP) from
// where it lies and compiles it unmodified against the ROS test doubles; oracle/Makefile puts the
// result in oracle/_ref/libref_global_costmap.so.  The node does all its work in its constructor
// (GC:5-18 -> setupRoadMap GC:52-61) and publishes once on /global_costmap; the harness constructs
// it on a CSV file in the edges.csv format and returns what it published.  No rasteriser
// arithmetic lives here.
//
// Memory: the reference holds the cost matrix as vector<vector<int>> (4 B per cell + 24 B per row)
// next to the int8 message: 0.7 GB at the launch-file size of 12000 x 12000.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdlib.h>
#include <string>
#include <vector>

#include "../include/prius_costmap.h"
#include "ros_doubles/ros_doubles.h"
#include "ros_doubles/tf/tf.h"
#include "ros_doubles/boost/algorithm/string.hpp"

#define private public
#include REF_GC_CPP
#undef private

extern "C" {

// Constructs Prius::GlobalCostmap on `csv_path`; copies the published grid (up to cap bytes) into
// out and returns the wall seconds of the constructor, or -1 if nothing was published.
double refg_build(const pg_params* p, const char* csv_path, int8_t* out, int64_t cap, int64_t* out_len) {
  auto& prm = ros_doubles::state().params;
  prm.clear();
  prm["local_costmap_res"] = p->res;
  prm["road_width"] = p->road_width;
  prm["num_lanes"] = p->num_lanes;
  prm["global_costmap_height"] = p->height_cells * p->res;
  prm["global_costmap_width"] = p->width_cells * p->res;
  auto& pub = ros_doubles::captured<nav_msgs::OccupancyGrid>();
  pub.clear();
  ros::NodeHandle nh, pnh("~");
  const auto t0 = std::chrono::steady_clock::now();
  {
    Prius::GlobalCostmap node(nh, pnh, csv_path);
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (pub.empty()) return -1.0;
  const auto& g = pub.back();
  if (out_len) *out_len = static_cast<int64_t>(g.data.size());
  const int64_t n = std::min<int64_t>(cap, static_cast<int64_t>(g.data.size()));
  if (out) std::memcpy(out, g.data.data(), n);
  pub.clear();
  pub.shrink_to_fit();
  return secs;
}

}  // extern "C"
