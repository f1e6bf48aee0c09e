// local_planner_b200.hpp -- host-side C++ mirror of the reference planner node's interface
// for the accelerated path, header-only, on top of the C ABI (include/prius_planner.h).
//
// The reference class Prius::LocalPlanner (local_planner.hpp:22-149) exposes nothing but
// three ROS callbacks and publishes three messages.  This class keeps the same method
// names, argument meaning and failure behaviour ("log and publish nothing") with the ROS
// message types replaced by plain structs of the same fields, so that
//   - ros/local_planner_b200_node.cpp is a field-by-field translation (it cannot be built
//     in this container: no ROS), and
//   - the same class runs headless in tests/cpp/test_host_node.cpp.
//
//   reference member                      here
//   costmapCallback        LP:658-661     costmapCallback(const OccupancyGrid&)
//   gazeboStatesCallback   LP:693-703     gazeboStatesCallback(const Twist&)
//   localNavCallback       LP:663-668     localNavCallback(const LocalNav&, const Transform2D&)
//   convertGoalToLocalFrame LP:670-691    (inside localNavCallback; the tf lookup is the caller's)
//   planPath               LP:25-82       pp_plan
//   publishGoal/Path/MPOutput LP:633-656  on_goal / on_path / on_mp std::function hooks
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "prius_planner.h"

namespace prius_b200 {

// nav_msgs/OccupancyGrid: info.{resolution,width,height} + data (row-major int8)
struct OccupancyGrid {
  double resolution = 0;
  uint32_t width = 0, height = 0;
  std::vector<int8_t> data;
};
// prius_msgs/LocalNav (prius_msgs/msg/LocalNav.msg:1-7)
struct LocalNav {
  double x = 0, y = 0, speed = 0, max_speed = 0, max_accel = 0;
};
// geometry_msgs/Twist.linear of the "prius" model (LP:693-703)
struct Twist {
  double linear_x = 0, linear_y = 0, linear_z = 0;
};
// prius_msgs/MotionPlanning (prius_msgs/msg/MotionPlanning.msg:1-7)
struct MotionPlanning {
  std::vector<double> yaw_rates, durations, speeds;
  double max_speed = 0, max_accel = 0;
};
// nav_msgs/Path in frame base_link: positions only (LP:530-538)
struct Path {
  std::vector<double> x, y;
};
// tf base_link <- map as a planar rigid transform (what lookupTransform("/base_link","/map")
// returns, LP:680): p_local = R(yaw) * p_map + t
struct Transform2D {
  double tx = 0, ty = 0, yaw = 0;
};

class LocalPlanner {
 public:
  // LocalPlanner::LocalPlanner (LP:5-18): parameters come from the launch file
  // (setupCostmap LP:390-399, getTreeParams LP:413-423) and the URDF (setupCollision LP:328-360)
  explicit LocalPlanner(const pp_params& params, int device = 0) : params_(params) {
    if (pp_create(&params_, device, &h_) != PP_OK) throw std::runtime_error(std::string("pp_create: ") + pp_last_error());
    path_xy_.resize(2 * kCapPath);
    yaw_rates_.resize(kCapPath);
    durations_.resize(kCapPath);
    speeds_.resize(kCapPath);
  }
  ~LocalPlanner() { pp_destroy(h_); }
  LocalPlanner(const LocalPlanner&) = delete;
  LocalPlanner& operator=(const LocalPlanner&) = delete;

  std::function<void(const MotionPlanning&)> on_mp;    // mp_pub,   "/mp"
  std::function<void(const Path&)> on_path;            // path_pub, "/local_path"
  std::function<void(double, double)> on_goal;         // goal_pub, "/goal"
  std::function<void(const std::string&)> on_error;    // ROS_ERROR_STREAM

  // LP:658-661 -> mapCostmapToCSpace (LP:425-435): the grid is copied, flipped on the device
  void costmapCallback(const OccupancyGrid& msg) {
    if (pp_set_grid(h_, msg.data.data(), static_cast<int32_t>(msg.width), static_cast<int32_t>(msg.height),
                    msg.resolution, PP_GRID_OCCUPANCY_MSG) != PP_OK)
      error(std::string("pp_set_grid: ") + pp_last_error());
    else
      have_grid_ = true;
  }

  // LP:693-703
  void gazeboStatesCallback(const Twist& prius_twist) { twist_ = prius_twist; }

  // LP:663-668: local_nav = *msg; convertGoalToLocalFrame(); planPath();
  // Returns the plan status (pp_plan_status) or -1 when nothing could be planned.
  int localNavCallback(const LocalNav& msg, const Transform2D& base_link_from_map) {
    LocalNav nav = msg;
    // LP:687-690: local_point = transform * global_point
    const double c = std::cos(base_link_from_map.yaw), s = std::sin(base_link_from_map.yaw);
    const double lx = c * msg.x - s * msg.y + base_link_from_map.tx;
    const double ly = s * msg.x + c * msg.y + base_link_from_map.ty;
    nav.x = lx;
    nav.y = ly;
    return planPath(nav);
  }

  // LP:25-82 + LP:84-96 (start state) + LP:503-583 (outputs)
  int planPath(const LocalNav& nav_local) {
    if (params_.collision_mode != PP_COLLISION_AS_SHIPPED && !have_grid_) {
      error("planPath: no costmap received yet");
      return -1;
    }
    pp_query q{};
    q.goal_x = nav_local.x;
    q.goal_y = nav_local.y;
    q.goal_speed = nav_local.speed;
    q.max_speed = nav_local.max_speed;
    q.max_accel = nav_local.max_accel;
    // LP:92: |prius_velocity.linear|
    q.start_speed = std::sqrt(twist_.linear_x * twist_.linear_x + twist_.linear_y * twist_.linear_y +
                              twist_.linear_z * twist_.linear_z);
    if (on_goal) on_goal(q.goal_x, q.goal_y);  // publishGoal (LP:89)
    pp_result r{};
    r.cap_path = kCapPath;
    r.path_xy = path_xy_.data();
    r.yaw_rates = yaw_rates_.data();
    r.durations = durations_.data();
    r.speeds = speeds_.data();
    const int rc = pp_plan(h_, &q, &r);
    last_ = r;
    if (rc != PP_OK && rc != PP_ERR_CAPACITY) {
      error(std::string("pp_plan: ") + pp_last_error());
      return -1;
    }
    if (r.status != PP_PLAN_FOUND) {
      // LP:34-39: "path plan time exceeded, attempting to replan" -- nothing is published
      error("path plan cap exceeded, attempting to replan");
      return r.status;
    }
    Path path;
    const int np = r.n_path < kCapPath ? r.n_path : kCapPath;
    for (int k = 0; k < np; ++k) {
      path.x.push_back(path_xy_[2 * k]);
      path.y.push_back(path_xy_[2 * k + 1]);
    }
    if (on_path) on_path(path);            // publishPath (LP:539)
    MotionPlanning mp;
    mp.max_accel = nav_local.max_accel;    // LP:546-547
    mp.max_speed = nav_local.max_speed;
    const int nm = r.n_profile < kCapPath ? r.n_profile : kCapPath;
    mp.yaw_rates.assign(yaw_rates_.begin(), yaw_rates_.begin() + nm);
    mp.durations.assign(durations_.begin(), durations_.begin() + nm);
    mp.speeds.assign(speeds_.begin(), speeds_.begin() + nm);
    if (on_mp) on_mp(mp);                  // publishMPOutput (LP:582)
    return r.status;
  }

  const pp_result& last_result() const { return last_; }
  pp_handle* handle() { return h_; }

 private:
  static constexpr int kCapPath = 4096;
  void error(const std::string& m) { if (on_error) on_error(m); }
  pp_params params_;
  pp_handle* h_ = nullptr;
  bool have_grid_ = false;
  Twist twist_;
  pp_result last_{};
  std::vector<double> path_xy_, yaw_rates_, durations_, speeds_;
};

}  // namespace prius_b200
