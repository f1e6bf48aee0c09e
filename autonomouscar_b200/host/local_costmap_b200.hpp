// local_costmap_b200.hpp -- host-side C++ mirror of the reference costmap node's interface
// (Prius::LocalCostmap, local_costmap.hpp:15-70) on top of the C ABI (include/prius_costmap.h),
// header-only.  Same callback names and "latest message wins" behaviour as the reference:
//
//   reference member                          here
//   rightLaserCallback / leftLaserCallback    LC:372-380   keep the latest scan
//   globalCostmapCallback                     LC:382-385   pc_set_global_grid (device copy)
//   centerLaserCallback -> pubOccGrid         LC:360-370   pc_build, then on_grid (occ_grid_pub, "/local_costmap")
//   listener.lookupTransform(...)             LC:89,99,164,217   the caller supplies CostmapTf (the ROS
//                                                               shim reads it from tf)
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "prius_costmap.h"

namespace prius_b200 {

// sensor_msgs/LaserScan: the fields mapPlanarScan reads (LC:109-113)
struct LaserScan {
  float angle_min = 0, angle_increment = 0, range_max = 0;
  std::vector<float> ranges;
};
// sensor_msgs/PointCloud.points as x, y, z triples
struct PointCloud {
  std::vector<float> xyz;
};
// what the node reads from tf during one calcOccGrid pass
struct PlanarTf {
  double roll = 0, tf_x = 0, tf_y = 0, z_map = 0, inv_x = 0, inv_y = 0;
};
struct CostmapTf {
  PlanarTf right, left;
  double center_z = 0;
  double map_from_base[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  double base_from_map[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
};
// the published nav_msgs/OccupancyGrid (frame base_link; origin / orientation are the shim's, LC:52-62)
struct CostmapGrid {
  double resolution = 0;
  uint32_t width = 0, height = 0;
  std::vector<int8_t> data;
};

class LocalCostmap {
 public:
  explicit LocalCostmap(const pc_params& params, int device = 0) : params_(params) {
    if (pc_create(&params_, device, &h_) != PP_OK) throw std::runtime_error(std::string("pc_create: ") + pp_last_error());
    grid_.resolution = params_.res;
    grid_.width = static_cast<uint32_t>(params_.width_cells);    // LC:62
    grid_.height = static_cast<uint32_t>(params_.height_cells);  // LC:61
    grid_.data.resize(static_cast<size_t>(params_.height_cells * params_.width_cells));
  }
  ~LocalCostmap() { pc_destroy(h_); }
  LocalCostmap(const LocalCostmap&) = delete;
  LocalCostmap& operator=(const LocalCostmap&) = delete;

  std::function<void(const CostmapGrid&)> on_grid;       // occ_grid_pub, "/local_costmap"
  std::function<void(const std::string&)> on_error;

  void rightLaserCallback(const LaserScan& msg) { right_ = msg; }   // LC:372-375
  void leftLaserCallback(const LaserScan& msg) { left_ = msg; }     // LC:377-380
  void globalCostmapCallback(const std::vector<int8_t>& data) {     // LC:382-385
    if (pc_set_global_grid(h_, data.data(), static_cast<int64_t>(data.size())) != PP_OK)
      error(std::string("pc_set_global_grid: ") + pp_last_error());
  }
  // LC:366-370: center_cloud = *msg; pubOccGrid();  Returns the pp_status of the build.
  int centerLaserCallback(const PointCloud& msg, const CostmapTf& tf) {
    pc_inputs in{};
    fill(&in.right, right_, tf.right);
    fill(&in.left, left_, tf.left);
    in.cloud_xyz = msg.xyz.empty() ? nullptr : msg.xyz.data();
    in.n_cloud = static_cast<int32_t>(msg.xyz.size() / 3);
    in.center_z = tf.center_z;
    for (int k = 0; k < 12; ++k) { in.map_from_base[k] = tf.map_from_base[k]; in.base_from_map[k] = tf.base_from_map[k]; }
    const int rc = pc_build(h_, &in, grid_.data.data());
    if (rc != PP_OK) { error(std::string("pc_build: ") + pp_last_error()); return rc; }
    if (on_grid) on_grid(grid_);
    return rc;
  }
  pc_handle* handle() { return h_; }

 private:
  static void fill(pc_planar_scan* s, const LaserScan& m, const PlanarTf& t) {
    s->ranges = m.ranges.empty() ? nullptr : m.ranges.data();
    s->n = static_cast<int32_t>(m.ranges.size());
    s->angle_min = m.angle_min; s->angle_increment = m.angle_increment; s->range_max = m.range_max;
    s->roll = t.roll; s->tf_x = t.tf_x; s->tf_y = t.tf_y; s->z_map = t.z_map; s->inv_x = t.inv_x; s->inv_y = t.inv_y;
  }
  void error(const std::string& m) { if (on_error) on_error(m); }
  pc_params params_;
  pc_handle* h_ = nullptr;
  LaserScan right_, left_;
  CostmapGrid grid_;
};

}  // namespace prius_b200
