// Runs ros/global_costmap_b200_node.cpp UNCHANGED against the ROS test doubles on a GPU box:
//   test_ros_global_costmap_shim <edges.csv> <out.bin> <res> <road_width> <num_lanes> <height_m> <width_m>
// writes the data of the /global_costmap message it publishes and prints its metadata.
#include <cstdio>
#include <fstream>
#include <string>

#define main shim_main
#include "../../ros/global_costmap_b200_node.cpp"
#undef main

int main(int argc, char** argv) {
  if (argc < 8) return 2;
  auto& prm = ros_doubles::state().params;
  prm["local_costmap_res"] = std::stod(argv[3]);
  prm["road_width"] = std::stod(argv[4]);
  prm["num_lanes"] = std::stod(argv[5]);
  prm["global_costmap_height"] = std::stod(argv[6]);
  prm["global_costmap_width"] = std::stod(argv[7]);
  ros_doubles::state().quiet = false;
  char* args[] = {argv[0], argv[1], nullptr};
  const int rc = shim_main(2, args);
  auto& pub = ros_doubles::captured<nav_msgs::OccupancyGrid>();
  if (rc != 0 || pub.empty()) { std::printf("published 0 rc %d\n", rc); return 1; }
  const auto& g = pub.back();
  std::ofstream o(argv[2], std::ios::binary);
  o.write(reinterpret_cast<const char*>(g.data.data()), g.data.size());
  std::printf("published %zu %s %u %u %a %a %a %zu\n", pub.size(), g.header.frame_id.c_str(), g.info.width, g.info.height,
              static_cast<double>(g.info.resolution), g.info.origin.position.x, g.info.origin.position.y, g.data.size());
  return 0;
}
