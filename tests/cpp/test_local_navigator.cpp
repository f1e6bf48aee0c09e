// Driver for tests/test_navigator_ref_pin.py: runs a scripted sequence of navigator events through the
// native LocalNavigator (autonomouscar_b200/host/route_rollout_b200.hpp) and prints what it publishes
// and holds after every loop pass, so the C++ mirror is checked against the reference module's outputs.
//   stdin:  "route n" x y ... | "odom x y qw qx qy qz" | "mask m" | "spin"
//   stdout: per spin: published(0/1) x y speed max_speed max_accel path_index obstacle F R L avoid_x avoid_y avoid_speed
#include <cstdio>
#include <iostream>
#include <string>

#include "route_rollout_b200.hpp"

int main() {
  prius_b200::LocalNavigator nav;
  std::string cmd;
  while (std::cin >> cmd) {
    if (cmd == "route") {
      int n;
      std::cin >> n;
      std::vector<prius_b200::RoutePoint> r(n);
      for (auto& p : r) std::cin >> p.first >> p.second;
      nav.on_path(r);
    } else if (cmd == "odom") {
      double v[6];
      for (double& d : v) std::cin >> d;
      nav.on_odometry(v[0], v[1], v[2], v[3], v[4], v[5]);
    } else if (cmd == "mask") {
      unsigned long m;
      std::cin >> m;
      nav.on_costmap(static_cast<uint32_t>(m));
    } else if (cmd == "spin") {
      prius_b200::LocalNavGoal g{0, 0, 0, 0, 0};
      const bool pub = nav.spin_once(&g);
      std::printf("%d %.17g %.17g %.17g %.17g %.17g %d %d", pub ? 1 : 0, g.x, g.y, g.speed, g.max_speed, g.max_accel,
                  nav.selector.path_index, nav.obstacle ? 1 : 0);
      for (int k = 0; k < 4; ++k) std::printf(" %d", nav.F[k] ? 1 : 0);
      for (int k = 0; k < 4; ++k) std::printf(" %d", nav.R[k] ? 1 : 0);
      for (int k = 0; k < 4; ++k) std::printf(" %d", nav.L[k] ? 1 : 0);
      std::printf(" %.17g %.17g %.17g\n", nav.avoid_x, nav.avoid_y, nav.avoid_speed);
    }
  }
  return 0;
}
