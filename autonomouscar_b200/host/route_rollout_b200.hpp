// route_rollout_b200.hpp -- closed-loop route roll-outs as native host
// code on top of the C ABI: the same loop as autonomouscar_b200/rollout.py (which documents the
// reference lines it mirrors), without an interpreter between the 20 Hz ticks.
//
//   densify_route     local_navigator.py:27-46      way-points every 1 m
//   WaypointSelector  local_navigator.py:586-609    closest way-point from the current index on, then forward
//                                                   until it is >= 20 m away
//   LocalNavigator    local_navigator.py:13-78, 259-368, 546-716   the node: callbacks + one loop pass, with the
//                                                   obstacle branch fed by the pc_front_regions mask (prius_costmap.h)
//   ProfileFollower   controller.cpp:144-158,178-185  cursor through the /mp profile
//   RouteRollout::tick  one LocalNav per car (goal in its base_link frame, LP:670-691), ONE pp_plan_batch call
//                       for all cars, then the profile is tracked ideally at the controller's 100 Hz
//                       (controller_node.cpp:14) -- the stand-in for Gazebo + the PIDs is this repository's.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "prius_planner.h"

namespace prius_b200 {

typedef std::pair<double, double> RoutePoint;

inline std::vector<RoutePoint> densify_route(const std::vector<RoutePoint>& path, double spacing = 1.0) {
  std::vector<RoutePoint> out;
  if (path.empty()) return out;
  out.push_back(path[0]);
  for (size_t i = 0; i + 1 < path.size(); ++i) {
    const double x0 = path[i].first, y0 = path[i].second, x1 = path[i + 1].first, y1 = path[i + 1].second;
    const double divide = std::sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0)) / spacing;
    if (divide > 1) {
      const double n = std::ceil(divide);
      const double dx = (x1 - x0) / n, dy = (y1 - y0) / n;
      for (int k = 1; k < static_cast<int>(n + 1); ++k) out.push_back(RoutePoint(x0 + k * dx, y0 + k * dy));
    } else {
      out.push_back(path[i + 1]);
    }
  }
  return out;
}

struct WaypointSelector {
  const std::vector<RoutePoint>* route = nullptr;
  double min_dist = 20.0;          // local_navigator.py:522
  int path_index = 0;
  RoutePoint select(double x, double y) {
    const std::vector<RoutePoint>& r = *route;
    const int n = static_cast<int>(r.size());
    auto dist = [&](int i) { return std::sqrt((x - r[i].first) * (x - r[i].first) + (y - r[i].second) * (y - r[i].second)); };
    double closest = 5000.0;
    for (int i = path_index; i < n; ++i) {            // :590-597
      const double d = dist(i);
      if (d < closest) { path_index = i; closest = d; } else break;
    }
    double d = dist(path_index);
    while (d < min_dist) {                            // :603-609
      if (path_index >= n - 1) break;                 // (the reference would run off the list here)
      path_index++;
      if (path_index == n - 1) break;
      d = dist(path_index);
    }
    return r[path_index];
  }
};

// One LocalNav message (prius_msgs/LocalNav)
struct LocalNavGoal { double x, y, speed, max_speed, max_accel; };

// The query generator node as its callbacks and one 20 Hz loop pass; on_costmap takes the region mask the
// device scan produced (pc_front_regions, bits PC_REGION_*) instead of the OccupancyGrid message.
struct LocalNavigator {
  WaypointSelector selector;
  bool got_path = false, got_pose = false, obstacle = false;
  bool F[4] = {false, false, false, false}, R[4] = {false, false, false, false}, L[4] = {false, false, false, false};
  double x = 0, y = 0, heading = 0;                   // currentCarPose (:76)
  double avoid_x = 0, avoid_y = 0, avoid_speed = 0;   // obstacle_avoid_wpx / _wpy / _spd (:498-500)
  std::vector<RoutePoint> dense;                      // owned copy when on_path() densifies

  void on_path(const std::vector<RoutePoint>& path) {               // :13-62, first message only
    if (got_path) return;
    dense = densify_route(path);
    selector.route = &dense;
    got_path = true;
  }
  void share_route(const std::vector<RoutePoint>* already_dense) { selector.route = already_dense; got_path = true; }
  void on_pose(double px, double py, double yaw) { x = px; y = py; heading = yaw; got_pose = true; }   // :66-78
  void on_odometry(double px, double py, double qw, double qx, double qy, double qz) {
    on_pose(px, py, std::atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz)));                 // :70-75
  }
  void on_costmap(uint32_t mask) {                                   // :81-368
    obstacle = (mask >> 12) & 1u;
    for (int k = 0; k < 4; ++k) {
      F[k] = (mask >> k) & 1u;
      if ((mask >> (16 + k)) & 1u) R[k] = (mask >> (4 + k)) & 1u;
      if ((mask >> (20 + k)) & 1u) L[k] = (mask >> (8 + k)) & 1u;
    }
    const double c = std::cos(heading), s = std::sin(heading);
    const double cl = std::cos(heading + 1.5708), sl = std::sin(heading + 1.5708);
    if (F[3]) { avoid_x = x + c * 22.5; avoid_y = y + s * 22.5; avoid_speed = 2.0; }                      // :278-290
    if (F[2]) { avoid_x = x + c * 15 + cl * 6; avoid_y = y + s * 15 + sl * 6; avoid_speed = 1.5; }        // :292-305
    if (F[1]) { avoid_x = x + c * 7.5 + cl * 7; avoid_y = y + s * 7.5 + sl * 7; avoid_speed = 1.0; }      // :307-320
    if (F[0]) { avoid_x = x + c * 5; avoid_y = y + s * 5; avoid_speed = 0.0; }                            // :322-335
  }
  // one loop pass (:551-716); false = nothing published
  bool spin_once(LocalNavGoal* out) {
    if (!got_pose) return false;
    got_pose = false;
    if (!got_path) return false;
    if (obstacle) { *out = LocalNavGoal{avoid_x, avoid_y, avoid_speed, 1.5, 1.5}; return true; }          // :561-583
    const RoutePoint wp = selector.select(x, y);                                                          // :586-624
    *out = LocalNavGoal{wp.first, wp.second, 1.5, 1.5, 1.5};
    return true;
  }
};

struct ProfileFollower {
  std::vector<double> yaw_rates, durations, speeds;
  int it = 0;
  double start = 0;
  void new_profile(double t, const double* yr, const double* du, const double* sp, int n) {   // motionPlanningCallback
    yaw_rates.assign(yr, yr + n); durations.assign(du, du + n); speeds.assign(sp, sp + n);
    it = 0; start = t;
  }
  bool command(double t, double* yaw_rate, double* speed) {      // controller.cpp:144-158
    if (it < static_cast<int>(durations.size()) && (t - start) >= durations[it] / 1000) { start = t; it++; }
    if (it >= static_cast<int>(yaw_rates.size())) return false;
    *yaw_rate = yaw_rates[it]; *speed = speeds[it];
    return true;
  }
};

class RouteRollout {
 public:
  struct Car { double x, y, yaw, v; };
  RouteRollout(pp_handle* planner, std::vector<RoutePoint> route, const std::vector<Car>& starts, int cap_path = 256)
      : h_(planner), route_(std::move(route)), cars_(starts), cap_(cap_path) {
    const size_t n = cars_.size();
    navigators_.resize(n); followers_.resize(n); queries_.resize(n); results_.resize(n);
    path_.resize(n * 2 * cap_); yr_.resize(n * cap_); du_.resize(n * cap_); sp_.resize(n * cap_);
    for (size_t k = 0; k < n; ++k) {
      navigators_[k].share_route(&route_);
      pp_result r{};
      r.cap_path = cap_;
      r.path_xy = &path_[k * 2 * cap_]; r.yaw_rates = &yr_[k * cap_]; r.durations = &du_[k * cap_]; r.speeds = &sp_[k * cap_];
      results_[k] = r;
    }
  }
  // one 20 Hz tick; returns the pp_status of the batch call.  masks (optional, one per car) = the region
  // mask of each car's current costmap (pc_front_regions): enables the navigator's obstacle branch.
  // banked: car k plans on grid k of the planner's grid bank (pp_set_grid_bank_dev) instead of the shared grid.
  int tick(const uint32_t* masks = nullptr, bool banked = false) {
    const int n = static_cast<int>(cars_.size());
    for (int k = 0; k < n; ++k) {
      const Car& c = cars_[k];
      LocalNavGoal wp{};
      navigators_[k].on_pose(c.x, c.y, c.yaw);
      if (masks) navigators_[k].on_costmap(masks[k]);
      navigators_[k].spin_once(&wp);
      const double cs = std::cos(c.yaw), sn = std::sin(c.yaw), dx = wp.x - c.x, dy = wp.y - c.y;
      pp_query q{};
      q.goal_x = cs * dx + sn * dy; q.goal_y = -sn * dx + cs * dy;            // LP:687-690
      q.goal_speed = wp.speed; q.max_speed = wp.max_speed; q.max_accel = wp.max_accel;   // local_navigator.py:524-527
      q.start_speed = c.v; q.query_id = static_cast<uint32_t>(k); q.reserved = static_cast<uint32_t>(k);
      queries_[k] = q;
    }
    const int rc = banked ? pp_plan_batch_banked(h_, n, queries_.data(), results_.data())
                          : pp_plan_batch(h_, n, queries_.data(), results_.data());
    if (rc != PP_OK && rc != PP_ERR_CAPACITY) return rc;
    plans += n;
    for (int k = 0; k < n; ++k) {
      const pp_result& r = results_[k];
      if (r.status != PP_PLAN_FOUND) continue;       // nothing published: the old profile stays (LP:34-39)
      found++;
      const int np = r.n_profile < cap_ ? r.n_profile : cap_;
      followers_[k].new_profile(t_, r.yaw_rates, r.durations, r.speeds, np);
    }
    const double dt = 1.0 / 100;                      // controller_node.cpp:14
    for (int sub = 0; sub < 5; ++sub) {               // 100 Hz / 20 Hz
      for (int k = 0; k < n; ++k) {
        double yaw_rate, speed;
        if (!followers_[k].command(t_, &yaw_rate, &speed)) continue;
        Car& c = cars_[k];
        c.yaw += yaw_rate * dt;
        c.v = speed;
        c.x += c.v * dt * std::cos(c.yaw);
        c.y += c.v * dt * std::sin(c.yaw);
      }
      t_ += dt;
    }
    return PP_OK;
  }
  const std::vector<Car>& cars() const { return cars_; }
  long long plans = 0, found = 0;

 private:
  pp_handle* h_;
  std::vector<RoutePoint> route_;
  std::vector<Car> cars_;
  int cap_;
  double t_ = 0;
  std::vector<LocalNavigator> navigators_;
  std::vector<ProfileFollower> followers_;
  std::vector<pp_query> queries_;
  std::vector<pp_result> results_;
  std::vector<double> path_, yr_, du_, sp_;
};

}  // namespace prius_b200
