// pp_near_spans.h -- host-side table behind the ORIENTED batch kernel's first test (GridDev::near4).  Plain C++ (no
// CUDA): pp_grid.cu uploads it once per handle; tests/test_near_spans.py compiles it with g++ and checks its one
// obligation -- every 4x4 block a lattice sample can land in is inside the span of its row -- against the oracle's
// own pose_to_fixed geometry.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace pp {

// Structuring elements of near4, per yaw bin:  which 4x4 blocks (di, dj) away from the pose's block can hold a
// lattice sample.  With the pose at continuous cell coordinates c inside its block and a sample at c + d (d in the
// footprint rectangle R(yaw), centred on the pose: half lengths du along (sin, cos) and dv along (cos, -sin) in
// (row, column) -- pose_to_fixed's u and v), the sample's block is (di, dj) away only if d lies in the open square
// 4 (di, dj) + (-4, 4)^2: a rectangle-against-square overlap test (separating axes: the square's two and the
// rectangle's two), taken over the yaws of the bin widened by 1e-3 rad (the kernel bins a float product).  The
// rectangle is inflated by 1.5 cells (the kernel's float cell index may be one off in each axis) plus what the
// yaw sampling step can move a corner.  Rows are stored as one span of columns {first, last} (first > last: empty).
// du, dv: half length / half width of the footprint in cells (Lattice: su * hu, sv * hv); n_bins: yaw bins over pi.
inline void near_span_table(double du, double dv, int n_bins, int* e_out, std::vector<int16_t>* table) {
  const int E = static_cast<int>(std::ceil((std::sqrt(du * du + dv * dv) + 1.5 + 4.0) / 4.0)) + 1;
  const int rows = 2 * E + 1, NS = 32;
  table->assign(static_cast<size_t>(n_bins) * rows * 2, 0);
  for (int b = 0; b < n_bins; ++b) {
    const double lo = b * 3.14159265358979323846 / n_bins - 1.0e-3, hi = (b + 1) * 3.14159265358979323846 / n_bins + 1.0e-3;
    const double pad = 1.5 + 0.05 + (du + dv) * (hi - lo) / NS;
    const double ha = du + pad, hb = dv + pad, hs = 4.0;
    std::vector<int> first(rows, 1), last(rows, 0);
    for (int k = 0; k <= NS; ++k) {
      const double t = lo + (hi - lo) * k / NS;
      const double ux = std::sin(t), uy = std::cos(t), vx = std::cos(t), vy = -std::sin(t);
      for (int di = -E; di <= E; ++di)
        for (int dj = -E; dj <= E; ++dj) {
          const double cx = 4.0 * di, cy = 4.0 * dj;
          if (std::fabs(cx) >= hs + ha * std::fabs(ux) + hb * std::fabs(vx)) continue;
          if (std::fabs(cy) >= hs + ha * std::fabs(uy) + hb * std::fabs(vy)) continue;
          if (std::fabs(cx * ux + cy * uy) >= ha + hs * (std::fabs(ux) + std::fabs(uy))) continue;
          if (std::fabs(cx * vx + cy * vy) >= hb + hs * (std::fabs(vx) + std::fabs(vy))) continue;
          int& f = first[di + E];
          int& l = last[di + E];
          if (f > l) { f = dj; l = dj; } else { f = std::min(f, dj); l = std::max(l, dj); }
        }
    }
    for (int r = 0; r < rows; ++r) {
      (*table)[(static_cast<size_t>(b) * rows + r) * 2] = static_cast<int16_t>(first[r]);
      (*table)[(static_cast<size_t>(b) * rows + r) * 2 + 1] = static_cast<int16_t>(last[r]);
    }
  }
  *e_out = E;
}

}  // namespace pp
