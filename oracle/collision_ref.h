// TEST INFRASTRUCTURE -- CPU oracle (see det_math_ref.h for the usage rule).
//
// Grid ingest and the three collision semantics, restated from
//   catkin_ws/src/local_planner/src/local_planner.cpp ("LP"):
//     mapCostmapToCSpace / calcGridLocation   LP:425-441
//     checkCollision                          LP:443-458
//     calcCollisionMatrix                     LP:467-496
//     interpolatePoints (edge spacing rule)   LP:301-316
// plus the ORIENTED extension, whose arithmetic is specified here and
// in DESIGN.md and must be reproduced bit for bit by the CUDA kernels.
// Pinned: the grid flip and the REF_AABB booleans equal the reference binary's (oracle/_ref,
// tests/test_ref_pin.py); ORIENTED has no reference counterpart and is pinned by its own spec.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "../include/prius_planner.h"
#include "det_math_ref.h"

namespace oracle {

// The planner's private copy of the costmap: m_c_space[i][j], i < width, j < height
// (local_planner.hpp:114, LP:401-411).  Stored flat, i-major.
struct CSpace {
  int width = 0;   // first index count  (m_local_costmap_width)
  int height = 0;  // second index count (m_local_costmap_height)
  double res = 0;
  std::vector<int8_t> cell;  // cell[i * height + j]
  int8_t at(int i, int j) const { return cell[static_cast<size_t>(i) * height + j]; }
};

// LP:425-441.  location = H*W - i*H - j - 1, computed in double and truncated (LP:439).
inline void cspace_from_occupancy_msg(const int8_t* data, int width, int height, double res,
                                      CSpace* out) {
  out->width = width;
  out->height = height;
  out->res = res;
  out->cell.assign(static_cast<size_t>(width) * height, 0);
  for (int i = 0; i < width; ++i) {
    for (int j = 0; j < height; ++j) {
      const double xi = i, yj = j;
      const int location = static_cast<int>(static_cast<double>(height * width) - xi * height - yj - 1);
      out->cell[static_cast<size_t>(i) * height + j] = data[location];
    }
  }
}

inline void cspace_from_cspace(const int8_t* data, int width, int height, double res, CSpace* out) {
  out->width = width;
  out->height = height;
  out->res = res;
  out->cell.assign(data, data + static_cast<size_t>(width) * height);
}

// double -> int conversions of the reference are plain C++ truncations (UB outside the
// int range).  Defined behaviour here and on the GPU: anything not strictly inside
// +-2^30 means "too far away to touch the grid".
inline bool trunc_to_int(double v, int* out) {
  if (!(std::fabs(v) < 1073741824.0)) return false;
  *out = static_cast<int>(v);
  return true;
}

// ---- REF_AABB: LP:446-457 over LP:467-496, yaw unused ------------------------------
// Quirks kept: the car LENGTH runs along the first (lateral) index; both double->int
// conversions truncate toward zero; cells outside the grid are skipped ("clipped cells
// ignored" -- the reference indexes its row vector out of bounds there).
inline bool collide_ref_aabb(const CSpace& g, double car_length, double car_width, double px,
                             double py) {
  int base_x, base_y, min_x, max_x, min_y, max_y;
  if (!trunc_to_int(g.width / 2 - py / g.res, &base_x)) return false;
  if (!trunc_to_int(g.height / 2 - px / g.res, &base_y)) return false;
  if (!trunc_to_int(base_x - car_length / 2 / g.res, &min_x)) return false;
  if (!trunc_to_int(base_x + car_length / 2 / g.res, &max_x)) return false;
  if (!trunc_to_int(base_y - car_width / 2 / g.res, &min_y)) return false;
  if (!trunc_to_int(base_y + car_width / 2 / g.res, &max_y)) return false;
  for (int x = min_x; x < max_x; ++x) {
    if (x < 0 || x > g.width - 1) continue;
    for (int y = min_y; y < max_y; ++y) {
      if (y < 0 || y > g.height - 1) continue;
      if (g.at(x, y) == 100) return true;
    }
  }
  return false;
}

// ---- ORIENTED (extension, row S3) ----------------------------------------------------
// The footprint is a lattice of nu x nv sample points spanning the car_length x car_width
// rectangle, rotated by yaw about the pose.  Points are evaluated in CELL coordinates as
// 32-bit fixed point (24 fractional bits) relative to an anchor cell, so every lattice
// point is an exact integer expression -- associative, hence identical for any parallel
// evaluation order.
struct Lattice {
  int nu = 0, nv = 0;        // points along the length / across the width
  double su = 0, sv = 0;     // lattice step in cells
  double hu = 0, hv = 0;     // (nu-1)/2, (nv-1)/2
};
constexpr int kAnchorOffset = 48;      // anchor = floor(centre) - 48 cells
constexpr int kFracBits = 24;
constexpr double kFixedOne = 16777216.0;  // 2^24

inline int ceil_tol(double q) { return static_cast<int>(std::ceil(q - 1e-9)); }

inline Lattice make_lattice(double car_length, double car_width, double res) {
  Lattice l;
  const double lu = car_length / res, lv = car_width / res;
  l.nu = ceil_tol(lu) + 1;
  l.nv = ceil_tol(lv) + 1;
  l.su = l.nu > 1 ? lu / (l.nu - 1) : 0.0;
  l.sv = l.nv > 1 ? lv / (l.nv - 1) : 0.0;
  l.hu = (l.nu - 1) * 0.5;
  l.hv = (l.nv - 1) * 0.5;
  return l;
}
// the kernels support footprints whose half diagonal (+1 cell) stays below 47 cells
inline bool lattice_supported(const Lattice& l) {
  const double du = l.su * l.hu, dv = l.sv * l.hv;
  return std::sqrt(du * du + dv * dv) + 1.0 < 47.0 && l.nu >= 1 && l.nv >= 1;
}

struct PoseFixed {
  bool valid = false;
  int ai = 0, aj = 0;              // anchor cell
  int32_t oi = 0, oj = 0;          // lattice origin (a=0,b=0), fixed point, relative to anchor
  int32_t ui = 0, uj = 0;          // step along the length
  int32_t vi = 0, vj = 0;          // step across the width
};

inline int32_t to_fixed(double v) { return static_cast<int32_t>(std::llrint(v * kFixedOne)); }

inline PoseFixed pose_to_fixed(const Lattice& l, int width, int height, double res, double x,
                               double y, double yaw) {
  PoseFixed p;
  double s, c;
  detm::sincos(yaw, &s, &c);
  const double ci = static_cast<double>(width / 2) - y / res;   // first index  <-> -y
  const double cj = static_cast<double>(height / 2) - x / res;  // second index <-> -x
  if (!(std::fabs(ci) < 1048576.0) || !(std::fabs(cj) < 1048576.0) || s != s) return p;
  p.ai = static_cast<int>(std::floor(ci)) - kAnchorOffset;
  p.aj = static_cast<int>(std::floor(cj)) - kAnchorOffset;
  const double ui = -(s * l.su), uj = -(c * l.su);
  const double vi = -(c * l.sv), vj = s * l.sv;
  const double oi = ((ci - static_cast<double>(p.ai)) - l.hu * ui) - l.hv * vi;
  const double oj = ((cj - static_cast<double>(p.aj)) - l.hu * uj) - l.hv * vj;
  p.oi = to_fixed(oi); p.oj = to_fixed(oj);
  p.ui = to_fixed(ui); p.uj = to_fixed(uj);
  p.vi = to_fixed(vi); p.vj = to_fixed(vj);
  p.valid = true;
  return p;
}

inline bool collide_oriented(const CSpace& g, const Lattice& l, double x, double y, double yaw) {
  const PoseFixed p = pose_to_fixed(l, g.width, g.height, g.res, x, y, yaw);
  if (!p.valid) return false;
  for (int a = 0; a < l.nu; ++a) {
    for (int b = 0; b < l.nv; ++b) {
      const int32_t fi = p.oi + a * p.ui + b * p.vi;
      const int32_t fj = p.oj + a * p.uj + b * p.vj;
      const int gi = p.ai + (fi >> kFracBits);
      const int gj = p.aj + (fj >> kFracBits);
      if (gi < 0 || gi >= g.width || gj < 0 || gj >= g.height) continue;
      if (g.at(gi, gj) == 100) return true;
    }
  }
  return false;
}

// world coordinates (metres) of the lattice points exactly as the fixed-point evaluation
// places them; used for the 1e-5 m footprint-transform tolerance test
inline int footprint_points(const Lattice& l, int width, int height, double res, double x, double y,
                            double yaw, double* out_xy, int capacity) {
  const PoseFixed p = pose_to_fixed(l, width, height, res, x, y, yaw);
  if (!p.valid) return 0;
  int n = 0;
  for (int a = 0; a < l.nu; ++a) {
    for (int b = 0; b < l.nv; ++b) {
      if (n >= capacity) return n;
      const int32_t fi = p.oi + a * p.ui + b * p.vi;
      const int32_t fj = p.oj + a * p.uj + b * p.vj;
      const double ci = static_cast<double>(p.ai) + static_cast<double>(fi) / kFixedOne;
      const double cj = static_cast<double>(p.aj) + static_cast<double>(fj) / kFixedOne;
      out_xy[2 * n + 0] = (static_cast<double>(height / 2) - cj) * res;
      out_xy[2 * n + 1] = (static_cast<double>(width / 2) - ci) * res;
      ++n;
    }
  }
  return n;
}

// one collision query under a mode (LP:443 signature: point + yaw)
inline bool collide_pose(const CSpace& g, const Lattice& l, int mode, double car_length,
                         double car_width, double x, double y, double yaw) {
  switch (mode) {
    case PP_COLLISION_AS_SHIPPED: return false;  // LP:445
    case PP_COLLISION_REF_AABB: return collide_ref_aabb(g, car_length, car_width, x, y);
    default: return collide_oriented(g, l, x, y, yaw);
  }
}

// along-edge test: the spacing rule of interpolatePoints (LP:308: n = int(dist / search_res * 2))
// applied BETWEEN parent and child; poses k = 1..max(n,1), the parent end is not re-tested
inline bool collide_edge(const CSpace& g, const Lattice& l, int mode, double car_length,
                         double car_width, double search_res, double x0, double y0, double x1,
                         double y1, double yaw) {
  const double dx = x1 - x0, dy = y1 - y0;
  const double dist = std::sqrt(dx * dx + dy * dy);
  int n;
  if (!trunc_to_int(dist / search_res * 2, &n)) return false;
  if (n < 1) n = 1;
  for (int k = 1; k <= n; ++k) {
    const double t = static_cast<double>(k) / static_cast<double>(n);
    const double px = x0 + t * dx, py = y0 + t * dy;
    if (collide_pose(g, l, mode, car_length, car_width, px, py, yaw)) return true;
  }
  return false;
}

}  // namespace oracle
