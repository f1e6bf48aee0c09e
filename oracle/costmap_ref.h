// TEST INFRASTRUCTURE -- CPU oracle (see det_math_ref.h for the usage rule).
//
// Restatement of the grid PRODUCER one step before the planner:
//   catkin_ws/src/local_costmap/src/local_costmap.cpp ("LC")
//     calcOccGrid LC:71-80 . clearMap LC:349-358 . mapPlanarScan LC:82-156 .
//     mapCenterScan LC:158-206 . inflateGrid LC:288-347 . mapRoadVectors LC:208-257 .
//     calcGridLocation LC:259-265 . calcGlobalGridLocation LC:267-275 . calcCartesianCoords LC:277-286
// with ROS / tf types replaced by plain numbers: every tf lookup of the reference becomes an
// input field (the ROS shim reads them from tf), so no quaternion arithmetic happens here.
//
// PINNED AGAINST THE REFERENCE ITSELF: oracle/_ref/libref_costmap.so is the reference's own
// local_costmap.cpp compiled unmodified against the ROS test doubles (ref_costmap_harness.cpp);
// tests/test_costmap_ref_pin.py shows this restatement (LibmMath policy) gives the same grid byte
// for byte.  The whole pass is order-independent (obstacle marks win over free marks win over
// unknown; inflation and the road overlay are max-reductions), which is what lets the CUDA path
// run it as parallel atomic-max passes.
//
// Defined behaviour where the reference has none: calcGridLocation results outside
// [0, W*H) are ignored wherever the reference indexes without a check (LC:131-151, 178-201,
// 250-254); double -> int conversions outside +-2^30 count as "outside".
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../include/prius_costmap.h"
#include "det_math_ref.h"
#include "planner_ref.h"   // LibmMath / DetMath, trunc_to_int

namespace oracle {

struct LibmMathC : LibmMath {
  static double tan(double x) { return std::tan(x); }
};
struct DetMathC : DetMath {
  static double tan(double x) { double s, c; detm::sincos(x, &s, &c); return s / c; }
};

template <class M>
class RefCostmap {
 public:
  explicit RefCostmap(const pc_params& p) : P(p) {
    len_ = static_cast<int>(P.height_cells * P.width_cells);   // LC:63
  }
  int length() const { return len_; }

  // LC:71-80
  void build(const pc_inputs& in, const int8_t* global_grid, int64_t global_len, int8_t* out) {
    global_grid_ = global_grid;
    global_len_ = global_len;
    occ_.assign(len_, -1);        // clearMap LC:349-358
    inflated_.assign(len_, -1);
    map_planar_scan(in.right, false);
    map_planar_scan(in.left, true);
    map_center_scan(in);
    inflate_grid();
    map_road_vectors(in);
    for (int k = 0; k < len_; ++k) out[k] = occ_[k];
  }

  // LC:259-265: members are doubles, the result is truncated twice
  bool grid_location(double x, double y, int* loc) const {
    int width_pos, height_pos, location;
    if (!trunc_to_int(x / P.res + P.width_cells / 2, &width_pos)) return false;
    if (!trunc_to_int(y / P.res + P.height_cells / 2, &height_pos)) return false;
    if (!trunc_to_int(width_pos + height_pos * P.width_cells - 1, &location)) return false;
    *loc = location;
    return true;
  }
  bool in_range(int loc) const { return loc >= 0 && loc < len_; }

 private:
  void mark_obstacle(int loc) { if (in_range(loc)) occ_[loc] = 100; }
  void mark_free(int loc) { if (in_range(loc) && occ_[loc] != 100) occ_[loc] = 0; }

  // LC:82-156
  void map_planar_scan(const pc_planar_scan& s, bool is_left) {
    if (!s.ranges || s.n <= 0) return;
    for (int i = 0; i < s.n; ++i) {
      const float r = s.ranges[i];
      if (r > s.range_max) continue;                                    // LC:111
      const float anglef = s.angle_min + static_cast<float>(i) * s.angle_increment;   // float arithmetic, LC:113
      const double angle = anglef;
      double x, y;
      if (!is_left) {                                                   // LC:115-119
        y = -(r * M::cos(angle) * M::cos(s.roll) - s.tf_x);
        x = r * M::sin(angle) * M::cos(s.roll) - s.tf_y;
      } else {                                                          // LC:121-125
        y = r * M::cos(angle) * M::cos(s.roll) - s.tf_x;
        x = -(r * M::sin(angle) * M::cos(s.roll) - s.tf_y);
      }
      int loc;
      const bool ok = grid_location(x, y, &loc);
      const double max_range = std::fabs(s.z_map / M::tan(s.roll));    // LC:127
      const double dist_from_laser = r;
      if (std::fabs(y) < max_range && dist_from_laser > P.obstacle_range) { if (ok) mark_obstacle(loc); }
      else if (ok) mark_free(loc);
      const double angle_to_base = M::atan2(y - s.inv_y, x - s.inv_x);  // LC:141
      const double dist_to_base = std::sqrt(M::sq(x - s.inv_x) + M::sq(y - s.inv_y));
      int n;
      if (!trunc_to_int(dist_to_base / P.res, &n)) n = 0;
      for (int k = 1; k < n; ++k) {                                     // LC:144-153
        const double x_ = x - k * P.res * M::cos(angle_to_base);
        const double y_ = y - k * P.res * M::sin(angle_to_base);
        int l2;
        if (grid_location(x_, y_, &l2)) mark_free(l2);
      }
    }
  }

  // LC:158-206
  void map_center_scan(const pc_inputs& in) {
    for (int p = 0; p < in.n_cloud; ++p) {
      const float px = in.cloud_xyz[3 * p], py = in.cloud_xyz[3 * p + 1], pz = in.cloud_xyz[3 * p + 2];
      int loc;
      const bool ok = grid_location(px, py, &loc);
      const double dist = std::sqrt(M::sq(px) + M::sq(py));
      const double thresh = in.center_z + P.z_ground_buffer;
      if (pz > thresh && dist > P.obstacle_range) { if (ok) mark_obstacle(loc); }
      else if (ok) mark_free(loc);
      const double angle_to_base = M::atan2(py, px);
      int n;
      if (!trunc_to_int(dist / P.res, &n)) n = 0;
      // LC:194 `if(!pt.z > ...)`: logical NOT binds first, so the test is (pt.z == 0 ? 1 : 0) > thresh
      const bool clear_ray = static_cast<double>(!pz) > thresh;
      if (!clear_ray) continue;
      for (int k = 1; k < n; ++k) {
        const double x = px - k * P.res * M::cos(angle_to_base);
        const double y = py - k * P.res * M::sin(angle_to_base);
        int l2;
        if (grid_location(x, y, &l2)) mark_free(l2);
      }
    }
  }

  // LC:288-347
  void inflate_grid() {
    std::vector<int> occupied;
    const int h_int = static_cast<int>(P.height_cells);
    const double max_dist = std::sqrt(M::sq(P.height_cells * P.res) + M::sq(P.width_cells * P.res));
    for (int count = 0; count < len_; ++count) {
      const int8_t pt = occ_[count];
      const std::div_t d = std::div(count, h_int);                      // LC:279-281
      const int height_pos = d.quot, width_pos = d.rem;
      const double x = (width_pos - P.width_cells / 2) * P.res;
      const double y = (height_pos - P.height_cells / 2) * P.res;
      const double dist_to_base = std::sqrt(M::sq(x) + M::sq(y));
      const double inflation_radius = M::sq(dist_to_base / max_dist) * P.max_inflation_radius;
      const double num = inflation_radius / P.res;
      const double angle_inc = 2 * M_PI / num;
      for (int i = 0; i < num; ++i) {
        const double angle = i * angle_inc;
        for (int j = 0; j < num; ++j) {
          const double x_ = x + j / num * inflation_radius * M::cos(angle);
          const double y_ = y + j / num * inflation_radius * M::sin(angle);
          int loc;
          if (!grid_location(x_, y_, &loc)) continue;
          if (loc >= 0 && loc < len_) {
            if (pt == 100) occupied.push_back(loc);
            if (pt == 0) inflated_[loc] = 0;
          }
        }
      }
    }
    for (int loc : occupied) inflated_[loc] = 100;
    for (int k = 0; k < len_; ++k)
      if (inflated_[k] == 100 || inflated_[k] == 0) occ_[k] = inflated_[k];
  }

  // tf::Transform applied to (x, y, 0): row . point + origin, left to right (Bullet's Vector3::dot)
  static void apply(const double m[12], double x, double y, double* ox, double* oy) {
    *ox = ((m[0] * x + m[1] * y) + m[2] * 0.0) + m[3];
    *oy = ((m[4] * x + m[5] * y) + m[6] * 0.0) + m[7];
  }

  // LC:208-257
  void map_road_vectors(const pc_inputs& in) {
    if (!global_grid_ || global_len_ < 1) return;                        // LC:210-213
    const double total_x = P.height_cells * P.res, total_y = P.width_cells * P.res;
    const double min_x = -total_x / 2, min_y = -total_y / 2;
    const int nx = static_cast<int>(P.height_cells * 2), ny = static_cast<int>(P.width_cells * 2);
    for (int i = 0; i < nx; ++i) {
      for (int j = 0; j < ny; ++j) {
        const double ux = min_x + i / static_cast<double>(nx) * total_x;
        const double uy = min_y + j / static_cast<double>(ny) * total_y;
        double gx, gy, lx, ly;
        apply(in.map_from_base, ux, uy, &gx, &gy);
        // calcGlobalGridLocation LC:267-275
        int cx, cy, wpos, hpos, gloc;
        if (!trunc_to_int(P.global_height_cells / 2 - gx / P.res, &cx)) continue;
        if (!trunc_to_int(P.global_width_cells / 2 - gy / P.res, &cy)) continue;
        if (!trunc_to_int(P.global_width_cells - cy, &wpos)) continue;
        if (!trunc_to_int(P.global_height_cells - cx, &hpos)) continue;
        if (!trunc_to_int(hpos + wpos * P.global_height_cells, &gloc)) continue;
        if (gloc < 0 || gloc >= global_len_) continue;
        apply(in.base_from_map, gx, gy, &lx, &ly);
        int lloc;
        if (!grid_location(lx, ly, &lloc) || !in_range(lloc)) continue;
        if (occ_[lloc] > global_grid_[gloc]) continue;                   // LC:250-254
        occ_[lloc] = global_grid_[gloc];
      }
    }
  }

  pc_params P;
  int len_;
  const int8_t* global_grid_ = nullptr;
  int64_t global_len_ = 0;
  std::vector<int8_t> occ_, inflated_;
};

}  // namespace oracle
