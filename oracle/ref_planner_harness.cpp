// TEST INFRASTRUCTURE -- C harness around the REFERENCE's own local planner.
//
// This translation unit #includes the reference's source file
//     /root/reference/catkin_ws/src/local_planner/src/local_planner.cpp       (REF_LP_CPP)
// from where it lies (nothing is copied into the repo) and compiles it against the ROS test
// doubles in oracle/ros_doubles/.  oracle/Makefile builds two shared objects into oracle/_ref/:
//   libref_planner_shipped.so  REF_VARIANT 0: the file exactly as shipped; checkCollision
//                              returns false at local_planner.cpp:445
//   libref_planner_aabb.so     REF_VARIANT 1: the same file streamed through
//                              `sed 445d` (drops only the `return false;` line) on its way to
//                              the compiler, which turns the reference's dead collision code
//                              (local_planner.cpp:446-457, 467-496) live.  No patched copy is
//                              ever written to disk.
// The harness drives the node the way ROS would -- costmapCallback, gazeboStatesCallback,
// localNavCallback -- captures what the node publishes and reads the tree out of the object.
// It contains no planner arithmetic of its own.
//
// Determinism: the reference aborts a plan after 1 s of ros::Time (local_planner.cpp:33-39).
// The test-double clock is a tick counter (one tick per ros::Time::now() call), so "1 s" is made
// to elapse exactly when `max_expansions` iterations of the LP:31 loop have completed or when
// open + closed nodes reach `max_nodes` -- the same two caps the oracle and the CUDA path use.
//
// Known undefined behaviour of the reference that the harness steers around (never "fixes"):
//   markVisited (LP:318-326) indexes without a bounds check -> keep queries inside the window;
//   calcCollisionMatrix (LP:479-493) indexes m_collision_matrix[x - min_x] after skipping the
//   rows with x < 0 -> refp_check_collision reports 2 ("undefined in the reference") instead of
//   calling it for poses whose box crosses the low-x border.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <queue>
#include <sstream>
#include <stdlib.h>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include "../include/prius_planner.h"
#include "ros_doubles/ros_doubles.h"
#include "ros_doubles/tf/tf.h"
#include "ros_doubles/urdf/model.h"

// the node's members are all private (local_planner.hpp:28-84); the harness has to reach the
// callbacks and the tree
#define private public
#include REF_LP_CPP
#undef private

#ifndef REF_VARIANT
#define REF_VARIANT 0
#endif

namespace {

using Prius::GraphNode;
using Prius::LocalPlanner;

struct RefPlanner {
  pp_params P;
  std::unique_ptr<LocalPlanner> node;
};

void install_params(const pp_params& p) {
  auto& prm = ros_doubles::state().params;
  prm.clear();
  prm["local_costmap_res"] = p.costmap_res;
  // the node divides metres by the resolution and truncates (LP:397-398); hand it metres that
  // give back exactly the requested cell counts
  auto metres = [&](int cells) {
    double m = cells * p.costmap_res;
    for (int k = 0; k < 8 && static_cast<int>(m / p.costmap_res) != cells; ++k)
      m = std::nextafter(m, static_cast<int>(m / p.costmap_res) < cells ? 1e300 : -1e300);
    return m;
  };
  prm["local_costmap_height"] = metres(p.costmap_height);
  prm["local_costmap_width"] = metres(p.costmap_width);
  prm["collision_buffer_distance"] = p.collision_buffer_distance;
  prm["time_step_ms"] = p.time_step_ms;
  prm["velocity_res"] = p.velocity_res;
  prm["heading_res"] = p.heading_res;
  prm["heading_diff"] = p.heading_diff;
  prm["goal_pos_tolerance"] = p.goal_pos_tolerance;
  prm["goal_speed_tolerance"] = p.goal_speed_tolerance;
  prm["search_res"] = p.search_res;
  // setupCollision (LP:328-360) reads the footprint as |origin.x| of two tf lookups
  tf::register_transform("front_right_middle_sonar_link", "back_right_middle_sonar_link",
                         tf::Transform(tf::Matrix3x3(), tf::Vector3(p.car_length, 0, 0)));
  tf::register_transform("rear_right_wheel", "rear_left_wheel",
                         tf::Transform(tf::Matrix3x3(), tf::Vector3(p.car_width, 0, 0)));
  tf::register_transform("rear_right_wheel", "center_laser_link",
                         tf::Transform(tf::Matrix3x3(), tf::Vector3(1.0, 0, 0)));
  // convertGoalToLocalFrame (LP:670-691): identity unless refp_set_base_from_map is called
  if (tf::registry().find("base_link<-map") == tf::registry().end())
    tf::register_transform("base_link", "map", tf::Transform());
}

RefPlanner* make(const pp_params* p) {
  install_params(*p);
  auto* r = new RefPlanner;
  r->P = *p;
  ros::NodeHandle nh, pnh("~");
  r->node.reset(new LocalPlanner(nh, pnh));
  r->node->m_max_velocity = p->max_velocity;   // local_planner.hpp:141 default member (15)
  return r;
}

void write_state(const GraphNode& n, double* s) {
  s[0] = n.child_point.first; s[1] = n.child_point.second;
  s[2] = n.parent_point.first; s[3] = n.parent_point.second;
  s[4] = n.heading; s[5] = n.velocity; s[6] = n.g; s[7] = n.cost;
}

}  // namespace

extern "C" {

struct refp_plan_out {
  // in
  int32_t cap_nodes, cap_path;
  // out
  int32_t status;        // 0 = the node published /mp and /local_path; 1 = it gave up (LP:34-39)
  int32_t n_open, n_closed, n_path, n_profile;
  int32_t n_loop_ticks;  // ros::Time::now() calls made by the LP:31 loop == iterations started
  double goal_x, goal_y; // what publishGoal put on /goal (after convertGoalToLocalFrame)
  double mp_max_speed, mp_max_accel;
  double* open_state;    // 8 doubles per entry of m_tree_open, in vector order
  double* closed_state;  // same for m_tree_closed
  double* path_xy;       // nav_msgs/Path poses
  double* yaw_rates; double* durations; double* speeds;   // prius_msgs/MotionPlanning
  uint8_t* visited;      // m_c_space_visited[i][j] != 0, width*height, or NULL
};

int refp_variant(void) { return REF_VARIANT; }

void* refp_create(const pp_params* p) {
  if (!p) return nullptr;
  try { return make(p); } catch (...) { return nullptr; }
}
void refp_destroy(void* h) { delete static_cast<RefPlanner*>(h); }

// planar map -> base_link transform used by convertGoalToLocalFrame (LP:680-690)
void refp_set_base_from_map(double x, double y, double yaw) {
  tf::Quaternion q; q.setRPY(0, 0, yaw);
  tf::register_transform("base_link", "map", tf::Transform(q, tf::Vector3(x, y, 0)));
}

// costmapCallback(msg) (LP:658-661) with msg->data = `data` (width*height int8, message order)
int refp_set_grid_msg(void* h, const int8_t* data, int32_t width, int32_t height) {
  auto* r = static_cast<RefPlanner*>(h);
  if (!r || !data || width != r->node->m_local_costmap_width || height != r->node->m_local_costmap_height) return 1;
  auto msg = std::make_shared<nav_msgs::OccupancyGrid>();
  msg->info.width = width; msg->info.height = height; msg->info.resolution = r->P.costmap_res;
  msg->data.assign(data, data + static_cast<size_t>(width) * height);
  r->node->costmapCallback(msg);
  return 0;
}
// m_c_space as out[i*H + j]
int refp_get_cspace(void* h, int8_t* out) {
  auto* r = static_cast<RefPlanner*>(h);
  const auto& g = r->node->m_c_space;
  for (size_t i = 0; i < g.size(); ++i)
    for (size_t j = 0; j < g[i].size(); ++j) out[i * g[i].size() + j] = static_cast<int8_t>(g[i][j]);
  return 0;
}
// footprint the node derived in setupCollision
int refp_footprint(void* h, double* length, double* width) {
  auto* r = static_cast<RefPlanner*>(h);
  *length = r->node->m_car_length; *width = r->node->m_car_width;
  return 0;
}

int refp_plan(void* h, const pp_query* q, refp_plan_out* o) {
  auto* r = static_cast<RefPlanner*>(h);
  if (!r || !q || !o) return 1;
  LocalPlanner& node = *r->node;
  // /gazebo/model_states -> prius_velocity (LP:693-703)
  auto ms = std::make_shared<gazebo_msgs::ModelStates>();
  ms->name = {"ground_plane", "prius"};
  ms->pose.resize(2); ms->twist.resize(2);
  ms->twist[1].linear.x = q->start_speed;
  node.gazeboStatesCallback(ms);
  // /local_nav_waypoints -> localNavCallback -> convertGoalToLocalFrame -> planPath (LP:663-668)
  auto nav = std::make_shared<prius_msgs::LocalNav>();
  nav->x = q->goal_x; nav->y = q->goal_y; nav->speed = q->goal_speed;
  nav->max_speed = q->max_speed; nav->max_accel = q->max_accel;

  ros_doubles::captured<prius_msgs::MotionPlanning>().clear();
  ros_doubles::captured<nav_msgs::Path>().clear();
  ros_doubles::captured<geometry_msgs::PoseStamped>().clear();
  ros_doubles::State& st = ros_doubles::state();
  st.ticks = 0;
  // loop iteration k (0-based) sees now() - start_time = (k + 1) ticks; "> 1 s" <=> k >= max_expansions
  st.ticks_per_sec = static_cast<double>(r->P.max_expansions);
  const size_t max_nodes = static_cast<size_t>(r->P.max_nodes);
  st.force_timeout = [&node, max_nodes]() {
    return node.m_tree_open.size() + node.m_tree_closed.size() >= max_nodes;
  };
  node.localNavCallback(nav);
  st.force_timeout = nullptr;

  const auto& mp = ros_doubles::captured<prius_msgs::MotionPlanning>();
  const auto& paths = ros_doubles::captured<nav_msgs::Path>();
  const auto& goals = ros_doubles::captured<geometry_msgs::PoseStamped>();
  o->status = mp.empty() ? 1 : 0;
  o->goal_x = goals.empty() ? 0 : goals[0].pose.position.x;
  o->goal_y = goals.empty() ? 0 : goals[0].pose.position.y;
  o->n_open = static_cast<int32_t>(node.m_tree_open.size());
  o->n_closed = static_cast<int32_t>(node.m_tree_closed.size());
  // ticks: publishGoal (1) + start_time (1) + one per loop iteration + the stamps of the outputs
  o->n_loop_ticks = 0;
  int rc = 0;
  if (o->n_open > o->cap_nodes || o->n_closed > o->cap_nodes) rc = 4;
  if (o->open_state)
    for (int k = 0; k < std::min(o->n_open, o->cap_nodes); ++k) write_state(node.m_tree_open[k], o->open_state + 8 * static_cast<size_t>(k));
  if (o->closed_state)
    for (int k = 0; k < std::min(o->n_closed, o->cap_nodes); ++k) write_state(node.m_tree_closed[k], o->closed_state + 8 * static_cast<size_t>(k));
  o->n_path = 0; o->n_profile = 0; o->mp_max_speed = 0; o->mp_max_accel = 0;
  if (!mp.empty()) {
    const auto& path = paths.back();
    o->n_path = static_cast<int32_t>(path.poses.size());
    if (o->n_path > o->cap_path) rc = 4;
    for (int k = 0; k < std::min(o->n_path, o->cap_path); ++k) {
      if (o->path_xy) { o->path_xy[2 * k] = path.poses[k].pose.position.x; o->path_xy[2 * k + 1] = path.poses[k].pose.position.y; }
    }
    const auto& m = mp.back();
    o->n_profile = static_cast<int32_t>(m.yaw_rates.size());
    o->mp_max_speed = m.max_speed; o->mp_max_accel = m.max_accel;
    for (int k = 0; k < std::min(o->n_profile, o->cap_path); ++k) {
      if (o->yaw_rates) o->yaw_rates[k] = m.yaw_rates[k];
      if (o->durations) o->durations[k] = m.durations[k];
      if (o->speeds) o->speeds[k] = m.speeds[k];
    }
  }
  if (o->visited) {
    const auto& v = node.m_c_space_visited;
    for (size_t i = 0; i < v.size(); ++i)
      for (size_t j = 0; j < v[i].size(); ++j) o->visited[i * v[i].size() + j] = v[i][j] != 0;
  }
  return rc;
}

// The reference's lattice generators, for known-answer tests (LP:191-232)
int refp_possible_velocities(void* h, const pp_query* q, double v, double* out, int cap) {
  auto* r = static_cast<RefPlanner*>(h);
  r->node->local_nav.speed = q->goal_speed; r->node->local_nav.max_speed = q->max_speed;
  r->node->local_nav.max_accel = q->max_accel;
  GraphNode n(point(0, 0), point(0, 0), 0, v);
  const auto vs = r->node->calcPossibleVelocities(n);
  for (size_t k = 0; k < vs.size() && static_cast<int>(k) < cap; ++k) out[k] = vs[k];
  return static_cast<int>(vs.size());
}
int refp_possible_yaws(void* h, double heading, double* out, int cap) {
  auto* r = static_cast<RefPlanner*>(h);
  GraphNode n(point(0, 0), point(0, 0), heading, 0);
  const auto ys = r->node->calcPossibleYaws(n);
  for (size_t k = 0; k < ys.size() && static_cast<int>(k) < cap; ++k) out[k] = ys[k];
  return static_cast<int>(ys.size());
}
// calcH / calcW / checkGoalDist of one node under one query (LP:115-144, 171-189)
int refp_costs(void* h, const pp_query* q, const double state[6], double* out_w, double* out_h, int* out_goal_dist) {
  auto* r = static_cast<RefPlanner*>(h);
  LocalPlanner& node = *r->node;
  node.local_nav.x = q->goal_x; node.local_nav.y = q->goal_y; node.local_nav.speed = q->goal_speed;
  node.local_nav.max_speed = q->max_speed; node.local_nav.max_accel = q->max_accel;
  node.markGoalPoint();
  GraphNode n(point(state[0], state[1]), point(state[2], state[3]), state[4], state[5]);
  *out_w = node.calcW(n);
  *out_h = node.calcH(n);
  *out_goal_dist = node.checkGoalDist(n) ? 1 : 0;
  return 0;
}

// checkCollision (LP:443-458) per pose: 0 / 1 = what the reference returns, 2 = the reference's
// behaviour is undefined for this pose (see the header comment).  In the shipped variant every
// answer is 0 because of LP:445.
double refp_check_collision(void* h, int64_t n, const float* poses, uint8_t* flags) {
  auto* r = static_cast<RefPlanner*>(h);
  LocalPlanner& node = *r->node;
  const auto t0 = std::chrono::steady_clock::now();
  for (int64_t k = 0; k < n; ++k) {
    const double x = poses[3 * k], y = poses[3 * k + 1], yaw = poses[3 * k + 2];
#if REF_VARIANT == 1
    // same expressions as LP:471-474, used ONLY to decide whether the call is defined
    const double bx = node.m_local_costmap_width / 2 - y / node.m_local_costmap_res;
    const double by = node.m_local_costmap_height / 2 - x / node.m_local_costmap_res;
    const double lo = bx - node.m_car_length / 2 / node.m_local_costmap_res;
    if (!(std::fabs(bx) < 1e9) || !(std::fabs(by) < 1e9) || lo < 1.0) { flags[k] = 2; continue; }
#endif
    flags[k] = node.checkCollision(point(x, y), yaw) ? 1 : 0;
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- a pool of node instances for the timed CPU arm (bench.py) -----------------------------
// Each instance owns its grid copy, as the node does (LP:425-435).  The reference node itself
// is single-threaded (local_planner_node.cpp:14-17); the pool runs one instance per thread on
// disjoint slices, which is the most host parallelism the reference's code can be given.
struct RefPool { std::vector<RefPlanner*> nodes; pp_params P; };

void* refp_pool_create(const pp_params* p, const int8_t* grid_msg, int n_threads) {
  if (!p || n_threads < 1) return nullptr;
  auto* pool = new RefPool;
  pool->P = *p;
  pool->nodes.resize(n_threads, nullptr);
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t]() {
      RefPlanner* r = make(p);
      if (grid_msg) refp_set_grid_msg(r, grid_msg, p->costmap_width, p->costmap_height);
      pool->nodes[t] = r;
    });
  for (auto& t : th) t.join();
  return pool;
}
void refp_pool_destroy(void* h) {
  auto* pool = static_cast<RefPool*>(h);
  if (!pool) return;
  for (RefPlanner* r : pool->nodes) delete r;
  delete pool;
}
// checkCollision over `n` poses, contiguous slice per instance; returns wall seconds
double refp_pool_check_collision(void* h, int64_t n, const float* poses, uint8_t* flags) {
  auto* pool = static_cast<RefPool*>(h);
  const int nt = static_cast<int>(pool->nodes.size());
  std::vector<std::thread> th;
  const auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&, t]() {
      const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
      refp_check_collision(pool->nodes[t], hi - lo, poses + 3 * lo, flags + lo);
    });
  for (auto& t : th) t.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// n independent LocalNav requests, one at a time per instance; returns wall seconds
double refp_pool_plan_batch(void* h, int32_t n, const pp_query* qs, int32_t* statuses, int32_t* n_nodes) {
  auto* pool = static_cast<RefPool*>(h);
  const int nt = static_cast<int>(pool->nodes.size());
  std::atomic<int> next{0};
  std::vector<std::thread> th;
  const auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&, t]() {
      for (;;) {
        const int k = next.fetch_add(1);
        if (k >= n) break;
        refp_plan_out o;
        std::memset(&o, 0, sizeof(o));
        refp_plan(pool->nodes[t], &qs[k], &o);
        if (statuses) statuses[k] = o.status;
        if (n_nodes) n_nodes[k] = o.n_open + o.n_closed;
      }
    });
  for (auto& t : th) t.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
