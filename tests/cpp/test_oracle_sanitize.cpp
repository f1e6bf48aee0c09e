// The oracle restatements under AddressSanitizer + UndefinedBehaviorSanitizer (the
// reference has real undefined behaviour -- unchecked grid indices, top() of an empty queue, an
// unsigned loop bound that underflows -- and the oracle must have DEFINED behaviour in all of those).
// Exercises exactly those corners; any sanitizer report makes the process exit non-zero.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../oracle/collision_ref.h"
#include "../../oracle/costmap_ref.h"
#include "../../oracle/global_costmap_ref.h"
#include "../../oracle/planner_ref.h"
#include "../../oracle/sampling_ref.h"

using namespace oracle;

static pp_params launch_params() {
  pp_params p;
  std::memset(&p, 0, sizeof(p));
  p.costmap_res = 0.25; p.costmap_width = 300; p.costmap_height = 300; p.collision_buffer_distance = 1;
  p.time_step_ms = 150; p.velocity_res = 0.1; p.heading_res = 0.005; p.heading_diff = 0.065;
  p.goal_pos_tolerance = 0.1; p.goal_speed_tolerance = 0.5; p.search_res = 0.25; p.max_velocity = 15;
  p.car_length = 4.7; p.car_width = 1.586; p.max_expansions = 300; p.max_nodes = 6000;
  p.rrt_samples_per_round = 32; p.rrt_max_samples = 2000; p.rrt_goal_bias = 0.05; p.rrt_goal_tolerance = 1.0;
  return p;
}

int main() {
  pp_params p = launch_params();
  CSpace g;
  g.width = 300; g.height = 300; g.res = 0.25;
  g.cell.assign(300 * 300, 0);
  for (int i = 0; i < 300; ++i) g.cell[i * 300 + 100] = 100;            // a wall 12.5 m ahead
  int found = 0, capped = 0, exhausted = 0;
  for (int mode = 0; mode < 3; ++mode) {
    p.collision_mode = mode;
    // goals inside, far outside the window (markVisited / box indices leave the grid) and at the origin (divide by ||goal|| = 0)
    const double goals[5][2] = {{8, 0}, {20, 3}, {500, -400}, {0, 0}, {-30, 30}};
    for (const auto& gl : goals) {
      pp_query q;
      std::memset(&q, 0, sizeof(q));
      q.goal_x = gl[0]; q.goal_y = gl[1]; q.goal_speed = 1.5; q.max_speed = 1.5; q.max_accel = 1.5; q.start_speed = 0.3;
      RefPlanner<LibmMath> a(p, &g);
      PlanOutput o;
      a.plan(q, &o);
      RefPlanner<DetMath> b(p, &g);
      PlanOutput o2;
      b.plan(q, &o2);
      found += o.status == PP_PLAN_FOUND; capped += o.status == PP_PLAN_CAP; exhausted += o.status == PP_PLAN_EXHAUSTED;
    }
  }
  // a start pose boxed in on all sides: the frontier runs empty (top() on an empty queue in the reference)
  {
    CSpace full = g;
    full.cell.assign(300 * 300, 100);
    p.collision_mode = PP_COLLISION_REF_AABB;
    pp_query q;
    std::memset(&q, 0, sizeof(q));
    q.goal_x = 10; q.goal_speed = 1.5; q.max_speed = 1.5; q.max_accel = 1.5;
    RefPlanner<LibmMath> a(p, &full);
    PlanOutput o;
    a.plan(q, &o);
    exhausted += o.status == PP_PLAN_EXHAUSTED;
  }
  // collision primitives at and beyond every border, NaN / inf poses
  {
    const Lattice lat = make_lattice(p.car_length, p.car_width, g.res);
    const double xs[] = {-1e9, -37.6, -37.5, 0, 37.4, 37.5, 1e9, 1e300, NAN, INFINITY};
    int hits = 0;
    for (double x : xs)
      for (double y : xs)
        for (int mode = 1; mode < 3; ++mode) hits += collide_pose(g, lat, mode, p.car_length, p.car_width, x, y, 0.7);
    if (hits < 0) return 3;
  }
  // local costmap: points far outside the window, NaN ranges, empty inputs, a global map that is too small
  {
    pc_params c;
    std::memset(&c, 0, sizeof(c));
    c.res = 0.25; c.height_cells = 300; c.width_cells = 300; c.z_ground_buffer = 0.075; c.obstacle_range = 2;
    c.max_inflation_radius = 5; c.global_height_cells = 400; c.global_width_cells = 400;
    std::vector<float> ranges = {1.f, 29.f, 500.f, NAN, INFINITY, 0.f, 12.f};
    std::vector<float> cloud = {5, 5, 0, 1e6f, -1e6f, 1, NAN, 2, 3, -37.5f, -37.5f, 0, 37.4f, 37.4f, -1, 0, 0, 0};
    pc_inputs in;
    std::memset(&in, 0, sizeof(in));
    for (pc_planar_scan* s : {&in.right, &in.left}) {
      s->ranges = ranges.data(); s->n = static_cast<int>(ranges.size()); s->angle_min = -2.2f; s->angle_increment = 0.7f;
      s->range_max = 1000.f; s->roll = 0.05; s->tf_x = 2.7; s->tf_y = 1.0; s->z_map = 0.5; s->inv_x = -2.7; s->inv_y = -1.06;
    }
    in.cloud_xyz = cloud.data(); in.n_cloud = static_cast<int>(cloud.size() / 3); in.center_z = -1.8;
    const double id[12] = {1, 0, 0, 1500, 0, 1, 0, -900, 0, 0, 1, 0};
    std::memcpy(in.map_from_base, id, sizeof(id));
    std::memcpy(in.base_from_map, id, sizeof(id));
    std::vector<int8_t> global(400 * 400, 50), out(300 * 300);
    RefCostmap<LibmMathC> a(c);
    a.build(in, global.data(), static_cast<int64_t>(global.size()), out.data());
    RefCostmap<DetMathC> b(c);
    b.build(in, nullptr, 0, out.data());
    pc_inputs none;
    std::memset(&none, 0, sizeof(none));
    b.build(none, nullptr, 0, out.data());
  }
  // global road map: polylines that leave the map on every side, single points, zero-length segments
  {
    pg_params gp;
    std::memset(&gp, 0, sizeof(gp));
    gp.res = 0.25; gp.road_width = 11.1; gp.num_lanes = 2; gp.height_cells = 400; gp.width_cells = 400;
    std::vector<float> xy = {-80, -80, 80, 80, 300, -300, 0, 0, 0, 0, 10, 10, -55, 49.9f, -49.9f, 55};
    std::vector<int32_t> off = {0, 3, 4, 6, 8};
    std::vector<int8_t> out(400 * 400);
    RefGlobalCostmap<LibmMathC> a(gp);
    a.build(xy.data(), off.data(), 4, out.data(), static_cast<int64_t>(out.size()));
    RefGlobalCostmap<DetMathC> b(gp);
    b.build(xy.data(), off.data(), 0, out.data(), static_cast<int64_t>(out.size()));
  }
  // sampling planner with a goal outside the grid
  {
    p.collision_mode = PP_COLLISION_ORIENTED; p.time_step_ms = 1000;
    pp_query q;
    std::memset(&q, 0, sizeof(q));
    q.goal_x = 60; q.goal_y = 5; q.goal_speed = 1.5; q.max_speed = 1.5; q.max_accel = 1.5; q.start_speed = 1; q.seed = 9;
    RrtOutput o;
    rrt_plan(p, g, q, &o);
  }
  std::printf("found %d capped %d exhausted %d\n", found, capped, exhausted);
  return (found >= 2 && capped >= 2 && exhausted >= 1) ? 0 : 4;
}
