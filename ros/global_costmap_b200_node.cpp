// global_costmap_b200_node.cpp -- drop-in replacement of the reference executable
// global_costmap_node (catkin_ws/src/global_costmap/src/global_costmap_node.cpp): same node name,
// parameters (global_costmap.cpp:8-15), command-line argument (the edges.csv path, launch file :68)
// and topic; the road map is rasterised by pg_build -> libprius_planner_b200.so and published once,
// like GlobalCostmap::setupRoadMap -> pubRoadmap (GC:52-61, 219-222).
//
// Compiled unchanged against the ROS test doubles and checked against the reference node in
// tests/test_gpu_ros_shim.py (no ROS in this repository's container).
#include <nav_msgs/OccupancyGrid.h>
#include <ros/ros.h>

#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "prius_costmap.h"

namespace {

// GC:25-50: one polyline per line, points separated by ',', coordinates by ' '; only tokens with
// exactly three fields count.  (The +5 / -5 shift of GC:42-43 is applied by pg_build.)
void read_edges(const std::string& path, std::vector<float>* xy, std::vector<int32_t>* offsets) {
  std::ifstream in(path);
  offsets->assign(1, 0);
  for (std::string line; std::getline(in, line);) {
    size_t a = 0;
    for (;;) {
      const size_t b = line.find(',', a);
      const std::string tok = line.substr(a, b == std::string::npos ? std::string::npos : b - a);
      std::vector<std::string> f;
      size_t c = 0;
      for (;;) {
        const size_t d = tok.find(' ', c);
        f.push_back(tok.substr(c, d == std::string::npos ? std::string::npos : d - c));
        if (d == std::string::npos) break;
        c = d + 1;
      }
      if (f.size() == 3) {
        xy->push_back(std::strtof(f[0].c_str(), nullptr));
        xy->push_back(std::strtof(f[1].c_str(), nullptr));
      }
      if (b == std::string::npos) break;
      a = b + 1;
    }
    offsets->push_back(static_cast<int32_t>(xy->size() / 2));
  }
}

}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "global_costmap_node");
  ros::NodeHandle nh, pnh("~");
  if (argc < 2) { ROS_ERROR("usage: global_costmap_b200_node <edges.csv>"); return 1; }
  pg_params p;
  pg_default_params(&p);
  double h_m = p.height_cells * p.res, w_m = p.width_cells * p.res;
  pnh.getParam("/local_costmap_node/local_costmap_res", p.res);   // GC:8
  pnh.getParam("road_width", p.road_width);                       // GC:9
  pnh.getParam("num_lanes", p.num_lanes);                         // GC:10
  pnh.getParam("global_costmap_height", h_m);                     // GC:12
  pnh.getParam("global_costmap_width", w_m);                      // GC:13
  p.height_cells = h_m / p.res;                                   // GC:14-15
  p.width_cells = w_m / p.res;
  ros::Publisher pub = nh.advertise<nav_msgs::OccupancyGrid>("/global_costmap", 10);   // GC:7

  std::vector<float> xy;
  std::vector<int32_t> offsets;
  read_edges(argv[1], &xy, &offsets);
  nav_msgs::OccupancyGrid roadmap;                                // createOccGrid, GC:63-72
  roadmap.header.frame_id = "map";
  roadmap.info.resolution = p.res;
  roadmap.info.height = p.height_cells;
  roadmap.info.width = p.width_cells;
  roadmap.info.origin.position.x = -p.width_cells / 8;
  roadmap.info.origin.position.y = -p.height_cells / 8;
  roadmap.info.origin.orientation.w = 1;
  roadmap.data.resize(static_cast<size_t>(p.height_cells * p.width_cells));
  int8_t* dev = nullptr;
  if (pg_build(&p, 0, xy.data(), offsets.data(), static_cast<int32_t>(offsets.size()) - 1, roadmap.data.data(), &dev) != PP_OK) {
    ROS_ERROR_STREAM("pg_build: " << pp_last_error());
    return 1;
  }
  pg_free(0, dev);
  pub.publish(roadmap);                                           // GC:219-222
  ros::spin();
  return 0;
}
