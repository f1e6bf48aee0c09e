// TEST INFRASTRUCTURE -- CPU oracle (see det_math_ref.h for the usage rule).
//
// Extension rows S1-S3 of SURVEY.md section 8(a): the reference planner has no random
// sampling and no nearest-neighbour query, so the semantics below are defined
// by this build (DESIGN.md "Sampling planner") and the CUDA kernels must match them bit
// for bit.  The motion model is the reference's forward step, LP:155-157.
//
// PARITY UNPINNED BY THE REFERENCE for these rows: there is no reference code, test or golden vector
// to pin a sampler, a nearest-neighbour query or an oriented footprint against.  They are pinned by
// their own specification, the Random123 known answers for Philox and hand-derived cases
// (tests/test_oracle_kat.py); everything the reference DOES contain is pinned against oracle/_ref.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "collision_ref.h"
#include "det_math_ref.h"
#include "philox_ref.h"
#include "planner_ref.h"

namespace oracle {

constexpr float kPiF = 3.14159274101257324219f;      // float(pi)
constexpr float kTwoPiF = 6.28318548202514648438f;   // float(2*pi)

// S1: pose k of a batch = f(seed, stream, index), x/y uniform in the box, yaw in [-pi, pi)
inline void sample_pose(uint64_t seed, uint32_t stream, uint64_t index, float x_lo, float x_hi,
                        float y_lo, float y_hi, float out[3]) {
  const uint32_t ctr[4] = {static_cast<uint32_t>(index), static_cast<uint32_t>(index >> 32), stream,
                           kTagPose};
  const uint32_t key[2] = {static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
  const Philox4 r = philox4x32_10(ctr, key);
  out[0] = std::fmaf(u01_from_bits(r.v[0]), x_hi - x_lo, x_lo);
  out[1] = std::fmaf(u01_from_bits(r.v[1]), y_hi - y_lo, y_lo);
  out[2] = std::fmaf(u01_from_bits(r.v[2]), kTwoPiF, -kPiF);
}

// S2: squared distance in float32 with a fixed evaluation order; ties -> lowest index
inline float dist2_f(float qx, float qy, float nx, float ny) {
  const float dx = qx - nx, dy = qy - ny;
  return std::fmaf(dy, dy, dx * dx);
}
inline int nearest_index(const float* node_xy, int64_t n_nodes, float qx, float qy, float* d2_out) {
  int best = -1;
  float best_d = 0;
  for (int64_t k = 0; k < n_nodes; ++k) {
    const float d = dist2_f(qx, qy, node_xy[2 * k], node_xy[2 * k + 1]);
    if (best < 0 || d < best_d) { best = static_cast<int>(k); best_d = d; }
  }
  if (d2_out) *d2_out = best_d;
  return best;
}
// S2 on large trees: the same query -- minimum of dist2_f, ties -> lowest index -- answered EXACTLY through a
// uniform bucket grid instead of a scan over all nodes, so that the oracle can replay runs with ~10^6 nodes
// (BASELINE configs[3]).  Independent of the CUDA path's structure: nodes are binned on insertion (cell index
// clamped to the grid, so far-away nodes sit in border cells), a query walks Chebyshev rings of cells around its
// own cell and stops once the best distance found is strictly below a conservative bound on anything outside the
// block examined so far.  tests/test_oracle_kat.py checks it against nearest_index on random and degenerate sets.
class BucketGridNN {
 public:
  BucketGridNN(float x_lo, float x_hi, float y_lo, float y_hi, int cells_per_side)
      : x_lo_(x_lo), y_lo_(y_lo), n_(cells_per_side > 0 ? cells_per_side : 1) {
    cx_ = (x_hi - x_lo) / n_;
    cy_ = (y_hi - y_lo) / n_;
    if (!(cx_ > 0)) cx_ = 1;
    if (!(cy_ > 0)) cy_ = 1;
    cells_.resize(static_cast<size_t>(n_) * n_);
    i0_ = j0_ = n_;
    i1_ = j1_ = -1;
  }
  void insert(int id, float x, float y) {
    const int i = cell_of(x, x_lo_, cx_), j = cell_of(y, y_lo_, cy_);
    cells_[static_cast<size_t>(i) * n_ + j].push_back(id);
    i0_ = std::min(i0_, i); i1_ = std::max(i1_, i);
    j0_ = std::min(j0_, j); j1_ = std::max(j1_, j);
  }
  // nearest of the inserted nodes (positions in node_xy); -1 if none
  int nearest(const float* node_xy, float qx, float qy, float* d2_out) const {
    if (i1_ < 0) return -1;
    const int qi = cell_of(qx, x_lo_, cx_), qj = cell_of(qy, y_lo_, cy_);
    int best = -1;
    float best_d = 0;
    // rings that cannot touch the occupied bounding box hold no node: start at the first that can, end at the last
    const int r_first = std::max(std::max(i0_ - qi, qi - i1_), std::max(std::max(j0_ - qj, qj - j1_), 0));
    const int r_last = std::max(std::max(qi - i0_, i1_ - qi), std::max(qj - j0_, j1_ - qj));
    for (int r = r_first; r <= r_last; ++r) {
      const int a0 = qi - r, a1 = qi + r, b0 = qj - r, b1 = qj + r;
      auto scan = [&](int i, int j) {
        for (int id : cells_[static_cast<size_t>(i) * n_ + j]) {
          const float d = dist2_f(qx, qy, node_xy[2 * id], node_xy[2 * id + 1]);
          if (best < 0 || d < best_d || (d == best_d && id < best)) { best = id; best_d = d; }
        }
      };
      for (int i = std::max(a0, i0_); i <= std::min(a1, i1_); ++i) {
        if (i == a0 || i == a1) {                              // the two full rows of the ring
          for (int j = std::max(b0, j0_); j <= std::min(b1, j1_); ++j) scan(i, j);
        } else {                                               // in between: only the ring's two end columns
          if (b0 >= j0_ && b0 <= j1_) scan(i, b0);
          if (b1 >= j0_ && b1 <= j1_) scan(i, b1);
        }
      }
      if (best >= 0) {
        // everything not yet examined lies outside the block of cells [a0, a1] x [b0, b1]; a side of the block
        // that reaches the grid border has nothing beyond it (border cells hold the clamped far-away nodes)
        const double inf = 1e300;
        const double dl = a0 > 0 ? static_cast<double>(qx) - (static_cast<double>(x_lo_) + static_cast<double>(a0) * cx_) : inf;
        const double dr = a1 < n_ - 1 ? (static_cast<double>(x_lo_) + static_cast<double>(a1 + 1) * cx_) - static_cast<double>(qx) : inf;
        const double db = b0 > 0 ? static_cast<double>(qy) - (static_cast<double>(y_lo_) + static_cast<double>(b0) * cy_) : inf;
        const double dt = b1 < n_ - 1 ? (static_cast<double>(y_lo_) + static_cast<double>(b1 + 1) * cy_) - static_cast<double>(qy) : inf;
        double lb = std::min(std::min(dl, dr), std::min(db, dt));
        // the binning arithmetic is float32: give the cell borders a generous slack before trusting the bound
        lb -= 1e-3 * (static_cast<double>(cx_) + static_cast<double>(cy_)) + 1e-3;
        if (lb > 0 && static_cast<double>(best_d) < lb * lb * (1.0 - 1e-5)) break;
      }
    }
    if (d2_out) *d2_out = best_d;
    return best;
  }

 private:
  int cell_of(float v, float lo, float c) const {
    const float t = (v - lo) / c;
    if (!(t > 0)) return 0;                 // (NaN / below the grid)
    if (t >= static_cast<float>(n_)) return n_ - 1;
    return static_cast<int>(t);
  }
  float x_lo_, y_lo_, cx_, cy_;
  int n_;
  int i0_, i1_, j0_, j1_;                   // bounding box of the occupied cells
  std::vector<std::vector<int>> cells_;
};

// A6 as a query: first node whose 6 key doubles all compare == (graph_node.hpp:24-30)
inline int find_exact(const double* keys, int64_t n_nodes, const double q[6]) {
  for (int64_t k = 0; k < n_nodes; ++k) {
    const double* e = keys + 6 * k;
    if (e[0] == q[0] && e[1] == q[1] && e[2] == q[2] && e[3] == q[3] && e[4] == q[4] && e[5] == q[5])
      return static_cast<int>(k);
  }
  return -1;
}

// Round-synchronous sampling planner.
//   node 0 = (start_x, start_y, start_heading, start_speed)
//   round r resolves samples s = r*K .. r*K+K-1 against the tree as of the END of round r-1:
//     draw   : Philox(ctr = (s_lo, s_hi, query_id, kTagRrt), key = seed) -> r0..r3
//              goal-biased if u01(r2) < float(goal_bias), else uniform over the grid extent
//     nearest: S2 over float32 copies of the node positions
//     steer  : heading change toward the sample clamped to +-heading_diff; speed
//              v' = min(v + max_accel*dt, max_speed); step (v + v')*dt  [LP:155-157]
//     edge   : collide_edge() with the handle's collision mode
//     append : survivors in sample order
//   stops at the first appended node (lowest sample index) inside the goal tolerance box.
struct RrtNode {
  double x, y, heading, velocity;
  double step;  // edge length from the parent
  int parent;
};
struct RrtOutput {
  int status = PP_PLAN_CAP;
  int n_rounds = 0;
  int n_samples = 0;     // samples drawn
  int n_collisions = 0;  // samples whose edge collided
  std::vector<RrtNode> nodes;
  std::vector<int> path_ids;  // start -> goal
};

inline void rrt_sample_box(const pp_params& P, float* x_lo, float* x_hi, float* y_lo, float* y_hi) {
  // cell j = H/2 - x/res in [0, H)  <=>  x in (H/2 - H, H/2] * res ; likewise y with W
  *x_lo = static_cast<float>((P.costmap_height / 2 - P.costmap_height) * P.costmap_res);
  *x_hi = static_cast<float>((P.costmap_height / 2) * P.costmap_res);
  *y_lo = static_cast<float>((P.costmap_width / 2 - P.costmap_width) * P.costmap_res);
  *y_hi = static_cast<float>((P.costmap_width / 2) * P.costmap_res);
}

inline void rrt_plan(const pp_params& P, const CSpace& G, const pp_query& Q, RrtOutput* out) {
  *out = RrtOutput();
  const Lattice lat = make_lattice(P.car_length, P.car_width, P.costmap_res);
  float x_lo, x_hi, y_lo, y_hi;
  rrt_sample_box(P, &x_lo, &x_hi, &y_lo, &y_hi);
  const float bias = static_cast<float>(P.rrt_goal_bias);
  const float gxf = static_cast<float>(Q.goal_x), gyf = static_cast<float>(Q.goal_y);
  const double dt_ms = P.time_step_ms;
  const uint32_t key[2] = {static_cast<uint32_t>(Q.seed), static_cast<uint32_t>(Q.seed >> 32)};
  const int K = P.rrt_samples_per_round > 0 ? P.rrt_samples_per_round : 1;

  std::vector<float> xyf;  // float2 copies for S2
  // trees beyond kGridFrom nodes answer S2 through the bucket grid (same result as the scan, see BucketGridNN)
  constexpr int64_t kGridFrom = 4096;
  BucketGridNN grid(x_lo, x_hi, y_lo, y_hi, 128);
  auto push_node = [&](const RrtNode& n) {
    grid.insert(static_cast<int>(out->nodes.size()), static_cast<float>(n.x), static_cast<float>(n.y));
    out->nodes.push_back(n);
    xyf.push_back(static_cast<float>(n.x));
    xyf.push_back(static_cast<float>(n.y));
  };
  push_node(RrtNode{Q.start_x, Q.start_y, Q.start_heading, Q.start_speed, 0.0, -1});

  int goal_id = -1;
  int64_t s = 0;
  while (goal_id < 0 && s < P.rrt_max_samples && static_cast<int>(out->nodes.size()) < P.max_nodes) {
    const int64_t n_prev = static_cast<int64_t>(out->nodes.size());
    std::vector<RrtNode> fresh;
    for (int k = 0; k < K && s < P.rrt_max_samples; ++k, ++s) {
      const uint32_t ctr[4] = {static_cast<uint32_t>(s), static_cast<uint32_t>(s >> 32), Q.query_id,
                               kTagRrt};
      const Philox4 r = philox4x32_10(ctr, key);
      float sx, sy;
      if (u01_from_bits(r.v[2]) < bias) {
        sx = gxf; sy = gyf;
      } else {
        sx = std::fmaf(u01_from_bits(r.v[0]), x_hi - x_lo, x_lo);
        sy = std::fmaf(u01_from_bits(r.v[1]), y_hi - y_lo, y_lo);
      }
      out->n_samples++;
      // (every inserted node belongs to the tree as of the end of the previous round: fresh ones are pushed after it)
      const int near = n_prev >= kGridFrom ? grid.nearest(xyf.data(), sx, sy, nullptr)
                                           : nearest_index(xyf.data(), n_prev, sx, sy, nullptr);
      const RrtNode& nn = out->nodes[near];
      const double th = detm::atan2(static_cast<double>(sy) - nn.y, static_cast<double>(sx) - nn.x);
      double d = detm::wrap_pi(th - nn.heading);
      if (d > P.heading_diff) d = P.heading_diff;
      if (d < -P.heading_diff) d = -P.heading_diff;
      const double yaw = nn.heading + d;
      double v_next = nn.velocity + Q.max_accel * dt_ms / 1000;
      if (v_next > Q.max_speed) v_next = Q.max_speed;
      const double sum_v = nn.velocity + v_next;
      const double step = sum_v * dt_ms / 1000;
      if (!(step > 0)) continue;
      double sn, cs;
      detm::sincos(yaw, &sn, &cs);
      const double nx = nn.x + step * cs, ny = nn.y + step * sn;
      if (collide_edge(G, lat, P.collision_mode, P.car_length, P.car_width, P.search_res, nn.x, nn.y, nx,
                       ny, yaw)) {
        out->n_collisions++;
        continue;
      }
      fresh.push_back(RrtNode{nx, ny, yaw, v_next, step, near});
    }
    out->n_rounds++;
    for (const RrtNode& n : fresh) {
      push_node(n);
      if (goal_id < 0 && std::fabs(n.x - Q.goal_x) <= P.rrt_goal_tolerance &&
          std::fabs(n.y - Q.goal_y) <= P.rrt_goal_tolerance)
        goal_id = static_cast<int>(out->nodes.size()) - 1;
    }
  }
  if (goal_id >= 0) {
    out->status = PP_PLAN_FOUND;
    std::vector<int> rev;
    for (int id = goal_id; id >= 0; id = out->nodes[id].parent) rev.push_back(id);
    out->path_ids.assign(rev.rbegin(), rev.rend());
  }
}

}  // namespace oracle
