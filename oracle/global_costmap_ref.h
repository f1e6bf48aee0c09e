// TEST INFRASTRUCTURE -- CPU oracle (see det_math_ref.h for the usage rule).
//
// Restatement of the static road-map rasteriser two steps before the planner (SURVEY.md section 8(f)
// row N3):  catkin_ws/src/global_costmap/src/global_costmap.cpp ("GC")
//   readEdges GC:25-50 (the +5 / -5 shift only; parsing is the caller's) . constructRoadmapMatrix
//   GC:193-205 . clearMap GC:74-81 . generateRoads GC:91-162 . interpolatePoints GC:164-184 .
//   calcRoadmapMatrixCoords GC:186-191 . calcRoadmap GC:207-217 . calcGridLocation GC:83-89
//
// The reference's `point` is std::pair<FLOAT, float> (global_costmap.hpp:9): every stored point
// is rounded to float, differences of stored coordinates are float subtractions, everything else
// is double.  That mixture is reproduced operation for operation.
//
// PINNED AGAINST THE REFERENCE ITSELF: oracle/_ref/libref_global_costmap.so is the reference's own
// global_costmap.cpp compiled unmodified against the ROS test doubles; tests/test_costmap_ref_pin.py
// shows this restatement (LibmMath policy) publishes the same /global_costmap byte for byte.
//
// Defined behaviour where the reference has none: calcRoadmap writes roadmap.data[location] for
// locations up to len + H (GC:213-214, the whole y = 0 column lands past the end of the vector);
// those writes are dropped.  Every update of the cost matrix is `if (m > v) m = v` with v >= 0,
// i.e. m = min(m, trunc(v)): order-independent, which is what the CUDA path relies on.
#pragma once
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

#include "../include/prius_costmap.h"
#include "costmap_ref.h"

namespace oracle {

template <class M>
class RefGlobalCostmap {
 public:
  typedef std::pair<float, float> fpoint;   // global_costmap.hpp:9

  explicit RefGlobalCostmap(const pg_params& p) : P(p) {
    H_ = static_cast<int>(P.height_cells);   // GC:195-196
    W_ = static_cast<int>(P.width_cells);
  }

  // polyline k = points [offsets[k], offsets[k+1]) of xy (raw file values, before the GC:42-43 shift)
  void build(const float* xy, const int32_t* offsets, int n_sections, int8_t* out, int64_t out_len) {
    matrix_.assign(static_cast<size_t>(H_) * W_, 99);                       // GC:193-205
    for (int64_t k = 0; k < out_len; ++k) out[k] = -1;                      // GC:74-81
    for (int s = 0; s < n_sections; ++s) {
      std::vector<fpoint> section;
      for (int k = offsets[s]; k < offsets[s + 1]; ++k)
        section.push_back(fpoint(xy[2 * k] + 5, xy[2 * k + 1] - 5));        // GC:42-43: float + int -> float
      for (size_t i = 0; i + 1 < section.size(); ++i) road_segment(section[i], section[i + 1]);   // GC:95-98
    }
    // GC:207-217
    for (int x = 0; x < H_; ++x)
      for (int y = 0; y < W_; ++y) {
        const fpoint pt(static_cast<float>(x), static_cast<float>(y));
        int width_pos, height_pos, location;
        if (!trunc_to_int(P.width_cells - pt.second, &width_pos)) continue;          // GC:85
        if (!trunc_to_int(P.height_cells - pt.first, &height_pos)) continue;         // GC:86
        if (!trunc_to_int(height_pos + width_pos * P.height_cells, &location)) continue;  // GC:87
        if (location < 0 || location >= out_len) continue;
        out[location] = static_cast<int8_t>(matrix_[static_cast<size_t>(x) * W_ + y]);
      }
  }

 private:
  // GC:164-184
  std::vector<fpoint> interpolate(fpoint start, const fpoint& end, bool road) const {
    std::vector<fpoint> pts;
    double d_x = start.first - end.first;      // float subtraction, widened
    double d_y = start.second - end.second;
    const double angle = M::atan2(d_y, d_x);
    if (road) {
      d_x += P.road_width * M::cos(angle);
      d_y += P.road_width * M::sin(angle);
    }
    const double dist = std::sqrt(M::sq(d_x) + M::sq(d_y));
    int n;
    if (!trunc_to_int(dist / P.res * 2, &n)) n = 0;
    for (int i = 0; i < n; ++i) {
      const double x = start.first + i / static_cast<double>(n) * dist * M::cos(angle);
      const double y = start.second + i / static_cast<double>(n) * dist * M::sin(angle);
      pts.push_back(fpoint(static_cast<float>(x), static_cast<float>(y)));
    }
    return pts;
  }
  // GC:126-141 / 142-157: one stroke from the lane centre towards a lane edge
  void stroke(const std::vector<fpoint>& pts) {
    int count = 0;
    const int num_pts = static_cast<int>(pts.size());
    for (const fpoint& pt : pts) {
      const double value = 99 - static_cast<double>(count) / static_cast<double>(num_pts) * 99;
      int xi, yi;
      if (!trunc_to_int(P.height_cells / 2 - pt.first / P.res, &xi)) continue;     // GC:188-189
      if (!trunc_to_int(P.width_cells / 2 - pt.second / P.res, &yi)) continue;
      const float mx = static_cast<float>(xi), my = static_cast<float>(yi);       // returned as point
      if (mx < 0 || mx > static_cast<float>(static_cast<size_t>(H_) - 1) || my < 0 ||
          my > static_cast<float>(static_cast<size_t>(W_) - 1))
        continue;                                                                  // GC:132-135
      int& m = matrix_[static_cast<size_t>(mx) * W_ + static_cast<size_t>(my)];
      if (m > value) m = static_cast<int>(value);                                  // GC:136-139
      count++;
    }
  }
  // GC:97-160 for one polyline segment
  void road_segment(const fpoint& start, const fpoint& end) {
    const std::vector<fpoint> pts = interpolate(start, end, true);
    const double d_x = start.first - end.first;
    const double d_y = start.second - end.second;
    const double angle = M::atan2(d_y, d_x);
    const double perp = angle + M_PI / 2;
    const double lane_width = P.road_width / P.num_lanes;
    for (const fpoint& pt : pts) {
      const double first_x = pt.first - (P.road_width / 2 + lane_width / 2) * M::cos(perp);
      const double first_y = pt.second - (P.road_width / 2 + lane_width / 2) * M::sin(perp);
      for (int j = 0; j < P.num_lanes; ++j) {
        const fpoint lc(static_cast<float>(first_x + j * lane_width * M::cos(perp)),
                        static_cast<float>(first_y + j * lane_width * M::sin(perp)));
        const fpoint bot(static_cast<float>(lc.first - lane_width / 2 * M::cos(perp)),
                         static_cast<float>(lc.second - lane_width / 2 * M::sin(perp)));
        const fpoint top(static_cast<float>(lc.first + lane_width / 2 * M::cos(perp)),
                         static_cast<float>(lc.second + lane_width / 2 * M::sin(perp)));
        stroke(interpolate(lc, bot, false));
        stroke(interpolate(lc, top, false));
      }
    }
  }

  pg_params P;
  int H_, W_;
  std::vector<int> matrix_;
};

}  // namespace oracle
