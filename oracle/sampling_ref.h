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
  auto push_node = [&](const RrtNode& n) {
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
      const int near = nearest_index(xyf.data(), n_prev, sx, sy, nullptr);
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
