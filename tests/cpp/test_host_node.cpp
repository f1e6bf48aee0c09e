// Headless exercise of the host-side mirror of the reference node
// (autonomouscar_b200/host/local_planner_b200.hpp) against the C ABI, on a GPU box.
// Prints one line of numbers that tests/test_gpu_host_node.py compares with the oracle.
#include <cstdio>
#include <vector>

#include "local_planner_b200.hpp"

int main() {
  pp_params p;
  pp_default_params(&p);
  p.collision_mode = PP_COLLISION_REF_AABB;
  prius_b200::LocalPlanner planner(p, 0);
  int n_mp = -1, n_path = -1, errors = 0;
  double last_speed = -1, gx = 0, gy = 0, path_end_x = 0, path_end_y = 0;
  planner.on_mp = [&](const prius_b200::MotionPlanning& m) {
    n_mp = static_cast<int>(m.speeds.size());
    if (n_mp > 0) last_speed = m.speeds.back();
  };
  planner.on_path = [&](const prius_b200::Path& path) {
    n_path = static_cast<int>(path.x.size());
    path_end_x = path.x.back();
    path_end_y = path.y.back();
  };
  planner.on_goal = [&](double x, double y) { gx = x; gy = y; };
  planner.on_error = [&](const std::string&) { errors++; };

  prius_b200::LocalNav nav;
  nav.x = 30.0; nav.y = 12.0; nav.speed = 1.5; nav.max_speed = 1.5; nav.max_accel = 1.5;
  prius_b200::Transform2D tr;   // base_link <- map
  tr.tx = -10.0; tr.ty = -12.0; tr.yaw = 0.0;
  // before any costmap: must refuse (REF_AABB needs the grid) and publish nothing
  const int rc0 = planner.localNavCallback(nav, tr);
  // an empty 300x300 costmap in wire order, with one obstacle row that the flip must move
  prius_b200::OccupancyGrid g;
  g.resolution = 0.25; g.width = 300; g.height = 300;
  g.data.assign(300 * 300, 0);
  planner.costmapCallback(g);
  prius_b200::Twist tw;
  tw.linear_x = 0.6; tw.linear_y = 0.8; tw.linear_z = 0.0;   // |v| = 1.0
  planner.gazeboStatesCallback(tw);
  const int rc1 = planner.localNavCallback(nav, tr);
  const pp_result& r = planner.last_result();
  std::printf("%d %d %d %d %d %.17g %.17g %.17g %.17g %.17g %d %d %d\n", rc0, rc1, errors, n_path, n_mp, gx, gy,
              path_end_x, path_end_y, last_speed, r.n_nodes, r.n_open, r.n_closed);
  return 0;
}
