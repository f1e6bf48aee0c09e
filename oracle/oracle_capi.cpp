// TEST INFRASTRUCTURE -- CPU oracle (see det_math_ref.h for the usage rule).
//
// extern "C" surface of the oracle, loaded with ctypes by tests/ and bench.py's
// cpu_baseline / --impl reference legs.  It mirrors the product's C ABI
// (include/prius_planner.h) closely enough that parity tests call both with the same
// structs and compare the outputs field by field.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/prius_planner.h"
#include "collision_ref.h"
#include "costmap_ref.h"
#include "global_costmap_ref.h"
#include "philox_ref.h"
#include "planner_ref.h"
#include "sampling_ref.h"

using namespace oracle;

namespace {

CSpace wrap_cspace(const pp_params* p, const int8_t* cspace) {
  CSpace g;
  if (cspace) cspace_from_cspace(cspace, p->costmap_width, p->costmap_height, p->costmap_res, &g);
  else {
    g.width = p->costmap_width; g.height = p->costmap_height; g.res = p->costmap_res;
    g.cell.assign(static_cast<size_t>(g.width) * g.height, 0);
  }
  return g;
}

int fill_result(const PlanOutput& o, pp_result* r) {
  r->status = o.status;
  r->n_expansions = o.n_expansions;
  r->n_nodes = static_cast<int>(o.slots.size());
  r->n_open = o.n_open;
  r->n_closed = o.n_closed;
  r->n_path = static_cast<int>(o.path.size());
  r->n_profile = static_cast<int>(o.yaw_rates.size());
  r->n_candidates = o.n_candidates;
  r->n_collisions = o.n_collisions;
  uint64_t digest = 0;
  for (size_t id = 0; id < o.slots.size(); ++id) {
    const TreeNode& n = o.slots[id];
    const double key[6] = {n.cx, n.cy, n.px, n.py, n.heading, n.velocity};
    digest += node_digest(static_cast<int>(id), key);
  }
  r->tree_digest = digest;
  int rc = PP_OK;
  const int nn = r->n_nodes;
  if ((r->node_state || r->node_parent || r->node_closed_seq) && nn > r->cap_nodes) rc = PP_ERR_CAPACITY;
  const int ncopy = nn < r->cap_nodes ? nn : r->cap_nodes;
  for (int id = 0; id < ncopy; ++id) {
    const TreeNode& n = o.slots[id];
    if (r->node_state) {
      double* s = r->node_state + 8 * static_cast<size_t>(id);
      s[0] = n.cx; s[1] = n.cy; s[2] = n.px; s[3] = n.py;
      s[4] = n.heading; s[5] = n.velocity; s[6] = n.g; s[7] = n.cost;
    }
    if (r->node_closed_seq) r->node_closed_seq[id] = o.closed_seq[id];
    if (r->node_parent) r->node_parent[id] = first_node_with_child(o, n.px, n.py);
  }
  if (r->expansion_ids) {
    if (o.n_expansions > r->cap_expansions) rc = PP_ERR_CAPACITY;
    const int ne = o.n_expansions < r->cap_expansions ? o.n_expansions : r->cap_expansions;
    for (int k = 0; k < ne; ++k) r->expansion_ids[k] = o.expansion_ids[k];
  }
  if (r->n_path > r->cap_path && (r->path_xy || r->path_node_ids)) rc = PP_ERR_CAPACITY;
  const int np = r->n_path < r->cap_path ? r->n_path : r->cap_path;
  for (int k = 0; k < np; ++k) {
    if (r->path_xy) { r->path_xy[2 * k] = o.path[k].first; r->path_xy[2 * k + 1] = o.path[k].second; }
    if (r->path_node_ids) r->path_node_ids[k] = first_node_with_child(o, o.path[k].first, o.path[k].second);
  }
  const int nm = r->n_profile < r->cap_path ? r->n_profile : r->cap_path;
  for (int k = 0; k < nm; ++k) {
    if (r->yaw_rates) r->yaw_rates[k] = o.yaw_rates[k];
    if (r->durations) r->durations[k] = o.durations[k];
    if (r->speeds) r->speeds[k] = o.speeds[k];
  }
  return rc;
}

template <class M>
int plan_impl(const pp_params* p, const CSpace& g, const pp_query* q, pp_result* r, uint8_t* visited) {
  RefPlanner<M> planner(*p, &g);
  PlanOutput o;
  planner.plan(*q, &o);
  if (visited) std::memcpy(visited, o.visited.data(), o.visited.size());
  return fill_result(o, r);
}

}  // namespace

extern "C" {

int orc_cspace_from_msg(const int8_t* data, int width, int height, int8_t* out) {
  CSpace g;
  cspace_from_occupancy_msg(data, width, height, 1.0, &g);
  std::memcpy(out, g.cell.data(), g.cell.size());
  return 0;
}

// math_mode: 0 = glibc (closest to the reference binary), 1 = deterministic (GPU parity)
int orc_plan(const pp_params* p, const int8_t* cspace, const pp_query* q, int math_mode,
             pp_result* r, uint8_t* visited_or_null) {
  const CSpace g = wrap_cspace(p, cspace);
  return math_mode == 0 ? plan_impl<LibmMath>(p, g, q, r, visited_or_null)
                        : plan_impl<DetMath>(p, g, q, r, visited_or_null);
}

// Timed batch of plans on `n_threads` std::threads (one independent query per thread at a
// time, the reference node itself is single-threaded: local_planner_node.cpp:14-17).
// Returns wall seconds; statuses/expansions are written if the arrays are non-null.
double orc_plan_batch_timed(const pp_params* p, const int8_t* cspace, int n, const pp_query* qs,
                            int math_mode, int n_threads, int32_t* statuses, int32_t* n_nodes,
                            uint64_t* digests) {
  const CSpace g = wrap_cspace(p, cspace);
  std::atomic<int> next{0};
  auto worker = [&]() {
    for (;;) {
      const int k = next.fetch_add(1);
      if (k >= n) return;
      pp_result r;
      std::memset(&r, 0, sizeof(r));
      if (math_mode == 0) plan_impl<LibmMath>(p, g, &qs[k], &r, nullptr);
      else plan_impl<DetMath>(p, g, &qs[k], &r, nullptr);
      if (statuses) statuses[k] = r.status;
      if (n_nodes) n_nodes[k] = r.n_nodes;
      if (digests) digests[k] = r.tree_digest;
    }
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (n_threads <= 1) worker();
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int orc_possible_velocities(const pp_params* p, const pp_query* q, double v, double* out, int cap) {
  RefPlanner<DetMath> pl(*p, nullptr);
  pl.set_query(*q);
  const auto v_ = pl.possible_velocities(v);
  for (size_t k = 0; k < v_.size() && static_cast<int>(k) < cap; ++k) out[k] = v_[k];
  return static_cast<int>(v_.size());
}
int orc_possible_yaws(const pp_params* p, double heading, double* out, int cap) {
  RefPlanner<DetMath> pl(*p, nullptr);
  const auto y = pl.possible_yaws(heading);
  for (size_t k = 0; k < y.size() && static_cast<int>(k) < cap; ++k) out[k] = y[k];
  return static_cast<int>(y.size());
}

// poses = n x (x, y, yaw) float32; mode from params; n_threads slices the batch
double orc_collide_batch(const pp_params* p, const int8_t* cspace, int64_t n, const float* poses,
                         uint8_t* flags, int n_threads) {
  const CSpace g = wrap_cspace(p, cspace);
  const Lattice lat = make_lattice(p->car_length, p->car_width, p->costmap_res);
  auto run = [&](int64_t lo, int64_t hi) {
    for (int64_t k = lo; k < hi; ++k)
      flags[k] = collide_pose(g, lat, p->collision_mode, p->car_length, p->car_width, poses[3 * k],
                              poses[3 * k + 1], poses[3 * k + 2]) ? 1 : 0;
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (n_threads <= 1) run(0, n);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(run, n * t / n_threads, n * (t + 1) / n_threads);
    for (auto& t : th) t.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int orc_collide_edges(const pp_params* p, const int8_t* cspace, int64_t n, const float* edges,
                      uint8_t* flags) {
  const CSpace g = wrap_cspace(p, cspace);
  const Lattice lat = make_lattice(p->car_length, p->car_width, p->costmap_res);
  for (int64_t k = 0; k < n; ++k) {
    const float* e = edges + 5 * k;
    flags[k] = collide_edge(g, lat, p->collision_mode, p->car_length, p->car_width, p->search_res, e[0],
                            e[1], e[2], e[3], e[4]) ? 1 : 0;
  }
  return 0;
}

int orc_footprint_points(const pp_params* p, double x, double y, double yaw, double* out_xy, int cap) {
  const Lattice lat = make_lattice(p->car_length, p->car_width, p->costmap_res);
  return footprint_points(lat, p->costmap_width, p->costmap_height, p->costmap_res, x, y, yaw, out_xy, cap);
}
int orc_lattice_dims(const pp_params* p, int* nu, int* nv, int* supported) {
  const Lattice lat = make_lattice(p->car_length, p->car_width, p->costmap_res);
  *nu = lat.nu; *nv = lat.nv; *supported = lattice_supported(lat) ? 1 : 0;
  return 0;
}

int orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const Philox4 r = philox4x32_10(ctr, key);
  for (int k = 0; k < 4; ++k) out[k] = r.v[k];
  return 0;
}
int orc_sample_poses(uint64_t seed, uint32_t stream, uint64_t first, int64_t n, float x_lo, float x_hi,
                     float y_lo, float y_hi, float* out) {
  for (int64_t k = 0; k < n; ++k) sample_pose(seed, stream, first + k, x_lo, x_hi, y_lo, y_hi, out + 3 * k);
  return 0;
}
int orc_nearest(int64_t n_nodes, const float* node_xy, int64_t n_query, const float* query_xy,
                int32_t* idx, float* d2) {
  for (int64_t k = 0; k < n_query; ++k) {
    float d = 0;
    idx[k] = nearest_index(node_xy, n_nodes, query_xy[2 * k], query_xy[2 * k + 1], &d);
    if (d2) d2[k] = d;
  }
  return 0;
}
int orc_find_exact(int64_t n_nodes, const double* keys, int64_t n_query, const double* q, int32_t* idx) {
  for (int64_t k = 0; k < n_query; ++k) idx[k] = find_exact(keys, n_nodes, q + 6 * k);
  return 0;
}

void orc_det_sincos(int64_t n, const double* x, double* s, double* c) {
  for (int64_t k = 0; k < n; ++k) detm::sincos(x[k], &s[k], &c[k]);
}
void orc_det_atan2(int64_t n, const double* y, const double* x, double* out) {
  for (int64_t k = 0; k < n; ++k) out[k] = detm::atan2(y[k], x[k]);
}
void orc_det_wrap_pi(int64_t n, const double* d, double* out) {
  for (int64_t k = 0; k < n; ++k) out[k] = detm::wrap_pi(d[k]);
}

// S2 through the bucket grid (test hook: must equal orc_nearest on any input)
void orc_nearest_grid(int64_t n_nodes, const float* node_xy, int64_t n_query, const float* query_xy, float x_lo, float x_hi,
                      float y_lo, float y_hi, int cells_per_side, int32_t* idx, float* d2) {
  BucketGridNN g(x_lo, x_hi, y_lo, y_hi, cells_per_side);
  for (int64_t k = 0; k < n_nodes; ++k) g.insert(static_cast<int>(k), node_xy[2 * k], node_xy[2 * k + 1]);
  for (int64_t q = 0; q < n_query; ++q) idx[q] = g.nearest(node_xy, query_xy[2 * q], query_xy[2 * q + 1], &d2[q]);
}

// sampling planner; result layout as the product's pp_rrt_plan:
//   node_state[8*id] = x y parent.x parent.y heading velocity step 0 ; node_parent = tree parent
int orc_rrt_plan(const pp_params* p, const int8_t* cspace, const pp_query* q, pp_result* r) {
  const CSpace g = wrap_cspace(p, cspace);
  RrtOutput o;
  rrt_plan(*p, g, *q, &o);
  r->status = o.status;
  r->n_expansions = o.n_rounds;
  r->n_nodes = static_cast<int>(o.nodes.size());
  r->n_open = r->n_nodes;
  r->n_closed = 0;
  r->n_candidates = o.n_samples;
  r->n_collisions = o.n_collisions;
  r->n_path = static_cast<int>(o.path_ids.size());
  r->n_profile = r->n_path > 1 ? r->n_path - 1 : 0;
  uint64_t digest = 0;
  int rc = PP_OK;
  for (int id = 0; id < r->n_nodes; ++id) {
    const RrtNode& n = o.nodes[id];
    const double px = n.parent >= 0 ? o.nodes[n.parent].x : n.x;
    const double py = n.parent >= 0 ? o.nodes[n.parent].y : n.y;
    const double key[6] = {n.x, n.y, px, py, n.heading, n.velocity};
    digest += node_digest(id, key);
    if (id < r->cap_nodes) {
      if (r->node_state) {
        double* s = r->node_state + 8 * static_cast<size_t>(id);
        s[0] = n.x; s[1] = n.y; s[2] = px; s[3] = py; s[4] = n.heading; s[5] = n.velocity;
        s[6] = n.step; s[7] = 0;
      }
      if (r->node_parent) r->node_parent[id] = n.parent;
      if (r->node_closed_seq) r->node_closed_seq[id] = -1;
    } else if (r->node_state || r->node_parent) rc = PP_ERR_CAPACITY;
  }
  r->tree_digest = digest;
  for (int k = 0; k < r->n_path && k < r->cap_path; ++k) {
    const RrtNode& n = o.nodes[o.path_ids[k]];
    if (r->path_xy) { r->path_xy[2 * k] = n.x; r->path_xy[2 * k + 1] = n.y; }
    if (r->path_node_ids) r->path_node_ids[k] = o.path_ids[k];
    if (k + 1 < r->n_path) {  // profile entry k: segment k -> k+1, in the form of LP:576-580
      const RrtNode& m = o.nodes[o.path_ids[k + 1]];
      if (r->yaw_rates) r->yaw_rates[k] = (m.heading - n.heading) / (p->time_step_ms / 1000);
      if (r->durations) r->durations[k] = p->time_step_ms;
      if (r->speeds) r->speeds[k] = m.velocity;
    }
  }
  if (r->n_path > r->cap_path && (r->path_xy || r->path_node_ids)) rc = PP_ERR_CAPACITY;
  return rc;
}

// local costmap (row N2): one calcOccGrid pass (LC:71-80).  Returns wall seconds of the pass.
double orc_costmap_build(const pc_params* p, const pc_inputs* in, const int8_t* global_grid, int64_t global_len,
                         int math_mode, int8_t* out) {
  const auto t0 = std::chrono::steady_clock::now();
  if (math_mode == 0) { RefCostmap<LibmMathC> c(*p); c.build(*in, global_grid, global_len, out); }
  else { RefCostmap<DetMathC> c(*p); c.build(*in, global_grid, global_len, out); }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// global road map (row N3): GlobalCostmap::setupRoadMap (GC:52-61).  Returns wall seconds.
double orc_global_costmap_build(const pg_params* p, const float* xy, const int32_t* offsets, int n_sections,
                                int math_mode, int8_t* out, int64_t out_len) {
  const auto t0 = std::chrono::steady_clock::now();
  if (math_mode == 0) { RefGlobalCostmap<LibmMathC> c(*p); c.build(xy, offsets, n_sections, out, out_len); }
  else { RefGlobalCostmap<DetMathC> c(*p); c.build(xy, offsets, n_sections, out, out_len); }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
