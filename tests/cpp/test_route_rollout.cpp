// Native closed-loop roll-out (autonomouscar_b200/host/route_rollout_b200.hpp) on a GPU box:
//   test_route_rollout <cars> <ticks>
// prints the wall seconds of the timed ticks, the plan counters and every car's final state (%a).
// tests/test_rollout.py compares the states with the Python loop around the same planner.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "route_rollout_b200.hpp"

int main(int argc, char** argv) {
  const int cars = argc > 1 ? std::atoi(argv[1]) : 64, ticks = argc > 2 ? std::atoi(argv[2]) : 10;
  pp_params p;
  pp_default_params(&p);
  p.max_nodes = 20000;
  pp_handle* h = nullptr;
  if (pp_create(&p, 0, &h) != PP_OK) { std::fprintf(stderr, "pp_create: %s\n", pp_last_error()); return 1; }
  // global_path_planner.py:101-102
  const double rx[12] = {26.4527, 39.5884, 52.7228, 69.0941, 83.7454, 105.948, 123.622, 142.292, 162.782, 172.996, 171.868, 171.274};
  const double ry[12] = {13.998, 14.5571, 14.5571, 15.1161, 15.6748, 16.14200, 16.142, 16.0263, 16.2719, 48.2004, 82.6559, 122.569};
  std::vector<prius_b200::RoutePoint> path;
  for (int k = 0; k < 12; ++k) path.push_back(prius_b200::RoutePoint(rx[k], ry[k]));
  std::vector<prius_b200::RouteRollout::Car> starts;
  const double step = argc > 3 ? std::atof(argv[3]) : 2.0;
  for (int k = 0; k < cars; ++k) starts.push_back({step * k, 15.0 + 0.05 * (k % 5), 0.02 * ((k % 7) - 3), 0.1 * (k % 12)});
  prius_b200::RouteRollout ro(h, prius_b200::densify_route(path), starts);
  const int warm = argc > 4 ? std::atoi(argv[4]) : 0;
  for (int t = 0; t < warm; ++t) ro.tick();
  const auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < ticks; ++t)
    if (ro.tick() != PP_OK) { std::fprintf(stderr, "tick: %s\n", pp_last_error()); return 1; }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("%.9f %lld %lld\n", secs, ro.plans, ro.found);
  for (const auto& c : ro.cars()) std::printf("%a %a %a %a\n", c.x, c.y, c.yaw, c.v);
  pp_destroy(h);
  return 0;
}
