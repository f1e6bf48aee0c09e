// TEST INFRASTRUCTURE -- CPU oracle (see det_math_ref.h for the usage rule).
//
// Headless restatement of the reference local planner's search,
//   catkin_ws/src/local_planner/src/local_planner.cpp ("LP") and
//   catkin_ws/src/local_planner/include/local_planner/{graph_node,local_planner}.hpp
// with ROS types replaced by plain structs and the 1 s ros::Time guard (LP:33-39) replaced
// by expansion / node caps.  Every function names the reference lines it follows.
//
// PINNED AGAINST THE REFERENCE ITSELF: oracle/_ref is the reference's own local_planner.cpp
// compiled unmodified against ROS test doubles (oracle/Makefile, ref_planner_harness.cpp);
// tests/test_ref_pin.py shows this restatement (LibmMath policy) equals it bit for bit -- open
// and closed lists, path, profile, visited bitmap -- live and against tests/golden/ref_planner.npz.
// (The reference ships no tests or golden vectors of its own.)
//
// Math policy `M`:
//   LibmMath -- glibc sin/cos/atan2/pow, i.e. what the reference binary would call;
//   DetMath  -- det_math_ref.h; the policy the GPU path reproduces bit for bit.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <utility>
#include <vector>

#include "collision_ref.h"
#include "det_math_ref.h"

namespace oracle {

struct LibmMath {
  static double sin(double x) { return std::sin(x); }
  static double cos(double x) { return std::cos(x); }
  static double atan2(double y, double x) { return std::atan2(y, x); }
  static double sq(double x) { return std::pow(x, 2); }  // the reference writes pow(x, 2)
};
struct DetMath {
  static double sin(double x) { return detm::sin(x); }
  static double cos(double x) { return detm::cos(x); }
  static double atan2(double y, double x) { return detm::atan2(y, x); }
  static double sq(double x) { return x * x; }
};

// The implicit double -> int conversion of an `int abs(int)` call (A18).  C++ leaves values outside the int
// range undefined; the build defines them (as everywhere else) as "too far away":
// anything not strictly inside +-2^30, NaN included, fails the tolerance test.
inline bool int_abs_within(double v, double tol) {
  if (!(std::fabs(v) < 1073741824.0)) return false;
  return std::abs(static_cast<int>(v)) <= tol;
}

// graph_node.hpp:7-31.  `id` (slot number = order of openNode calls) and `closed_seq` are
// book-keeping added here; they take no part in equality or ordering.
struct TreeNode {
  double cx, cy;   // child_point
  double px, py;   // parent_point
  double heading;
  double velocity;
  double g = 0;
  double cost = 0;  // uninitialised in the reference until assigned
  int id = -1;
};
// graph_node.hpp:24-30 (std::pair == compares .first then .second)
inline bool same_node(const TreeNode& a, const TreeNode& b) {
  return a.cx == b.cx && a.cy == b.cy && a.px == b.px && a.py == b.py && a.heading == b.heading &&
         a.velocity == b.velocity;
}
// local_planner.hpp:103-109
struct CheaperCost {
  bool operator()(const TreeNode& a, const TreeNode& b) const { return a.cost > b.cost; }
};

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint64_t dbits(double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; }
// digest contribution of slot `id` with the 6 key doubles (sum over slots, wrapping)
inline uint64_t node_digest(int id, const double key[6]) {
  uint64_t h = static_cast<uint64_t>(id) + 1;
  for (int k = 0; k < 6; ++k) h = splitmix64(h ^ dbits(key[k]));
  return h;
}

struct PlanOutput {
  int status = PP_PLAN_CAP;
  int n_expansions = 0;
  int n_candidates = 0;
  int n_collisions = 0;
  std::vector<TreeNode> slots;          // by id; final state of every slot
  std::vector<int> closed_seq;          // by id; -1 = open
  std::vector<int> expansion_ids;
  std::vector<std::pair<double, double>> path;  // start -> goal
  std::vector<double> yaw_rates, durations, speeds;
  int n_open = 0, n_closed = 0;
  std::vector<uint8_t> visited;         // optional debug bitmap (A12), width*height
};

template <class M>
class RefPlanner {
 public:
  RefPlanner(const pp_params& p, const CSpace* grid) : P(p), G(grid) {
    lattice_ = make_lattice(p.car_length, p.car_width, p.costmap_res);
  }

  // LP:25-82
  void plan(const pp_query& q, PlanOutput* out) {
    Q = q;
    O = out;
    *O = PlanOutput();
    initialize();
    while (true) {
      // replaces LP:33-39
      if (O->n_expansions >= P.max_expansions || static_cast<int>(O->slots.size()) >= P.max_nodes) {
        finish(PP_PLAN_CAP);
        return;
      }
      if (frontier_.empty()) { finish(PP_PLAN_EXHAUSTED); return; }  // LP:40 top() on empty
      TreeNode current = frontier_.top();                            // LP:40
      const int closed_id = close_top();                             // LP:41
      O->expansion_ids.push_back(closed_id);
      O->n_expansions++;
      TreeNode goal_node;
      if (check_for_goal(current, &goal_node)) {                     // LP:42-43
        open_node(goal_node);                                        // LP:45
        const bool ok = calc_path(goal_node.cx, goal_node.cy);       // LP:47
        if (!ok) { finish(PP_PLAN_BROKEN); return; }
        calc_profile();                                              // LP:48
        finish(PP_PLAN_FOUND);
        return;
      }
      std::vector<TreeNode> neighbors = get_neighbors(current);      // LP:52
      for (TreeNode& nb : neighbors) {                               // LP:53-73
        mark_visited(nb.cx, nb.cy);                                  // LP:55
        const int open_pos = find_in(open_, nb);                     // LP:56
        const int closed_pos = find_in(closed_, nb);                 // LP:57
        if (closed_pos < 0) {
          if (open_pos < 0) {
            open_node(nb);                                           // LP:62
          } else {
            if (nb.g < open_[open_pos].g) {                          // LP:66-69
              const int keep = open_[open_pos].id;
              open_[open_pos] = nb;
              open_[open_pos].id = keep;
            }
            open_node(nb);                                           // LP:70
          }
        }
      }
      if (frontier_.empty()) { finish(PP_PLAN_EXHAUSTED); return; }
      close_top();                                                   // LP:80
    }
  }

  // exposed for unit tests ------------------------------------------------------------
  // LP:191-216
  std::vector<double> possible_velocities(double node_velocity) const {
    std::vector<double> v;
    const double time_step = P.time_step_ms / 1000;
    const double hi = node_velocity + Q.max_accel * time_step;
    const double lo = node_velocity - Q.max_accel * time_step;
    const double range = hi - lo;
    int n;
    if (!trunc_to_int(range / P.velocity_res, &n)) n = 0;
    for (int i = 0; i < n; ++i) {
      const double vel = lo + i * P.velocity_res;
      if (vel == 0 && Q.goal_speed != 0) continue;
      if (vel > Q.max_speed || vel < 0) continue;
      v.push_back(vel);
    }
    return v;
  }
  // LP:218-232 -- the loop bound is a DOUBLE (26 or 27 headings depending on rounding)
  std::vector<double> possible_yaws(double heading) const {
    std::vector<double> y;
    const double hi = heading + P.heading_diff;
    const double lo = heading - P.heading_diff;
    const double range = hi - lo;
    const double n = range / P.heading_res;
    for (int i = 0; i < n && i < kMaxYaws; ++i) y.push_back(lo + i * P.heading_res);
    return y;
  }
  void set_query(const pp_query& q) { Q = q; }
  static constexpr int kMaxYaws = 4096;  // guard against absurd parameters only

  // LP:115-137
  double calc_h(const TreeNode& n) const {
    double speed = n.velocity;
    if (speed == 0) speed = 0.01;
    const double gx = Q.goal_x, gy = Q.goal_y;
    const double max_dist = std::sqrt(M::sq(gx) + M::sq(gy));
    const double dist_to_goal = std::sqrt(M::sq(n.cx - gx) + M::sq(n.cy - gy));
    const double dist_h = dist_to_goal / max_dist * 100;
    const double yaw_to_goal = M::atan2(gy - n.cy, gx - n.cx);
    const double yaw_h = std::fabs(n.heading - yaw_to_goal) / (2 * M_PI) * 100;
    double speed_h = 100 - speed / P.max_velocity * 100;
    if (check_goal_dist(n)) speed_h = std::fabs(speed - Q.goal_speed) * 100;
    return std::fabs(dist_h + yaw_h + speed_h);
  }
  // LP:171-189
  bool check_goal_dist(const TreeNode& n) const {
    double next_vel = n.velocity + Q.max_accel * P.time_step_ms / 1000;
    if (next_vel > Q.max_speed) next_vel = Q.max_speed;
    const double avg = (next_vel + n.velocity) / 2;
    const double x = n.cx + avg * P.time_step_ms / 1000 * M::cos(n.heading);
    const double y = n.cy + avg * P.time_step_ms / 1000 * M::sin(n.heading);
    const double dx = x - Q.goal_x, dy = y - Q.goal_y;
    const double dist_to_goal = std::sqrt(M::sq(dx) + M::sq(dy));
    const double speed = n.velocity;
    const double t_decel = (speed - Q.goal_speed) / Q.max_accel;
    const double d_decel = speed * t_decel - Q.max_accel * M::sq(t_decel) / 2;
    return dist_to_goal + P.goal_pos_tolerance <= d_decel;
  }
  // LP:146-169
  std::vector<TreeNode> get_neighbors(const TreeNode& node) {
    const std::vector<double> vels = possible_velocities(node.velocity);
    const std::vector<double> yaws = possible_yaws(node.heading);
    std::vector<TreeNode> out;
    for (double vel : vels) {
      for (double yaw : yaws) {
        const double sum_v = node.velocity + vel;  // LP:155: a sum, named "average"
        const double x = node.cx + (sum_v * P.time_step_ms / 1000) * M::cos(yaw);
        const double y = node.cy + (sum_v * P.time_step_ms / 1000) * M::sin(yaw);
        O->n_candidates++;
        if (collide_pose(*G, lattice_, P.collision_mode, P.car_length, P.car_width, x, y, yaw)) {
          O->n_collisions++;
          continue;
        }
        TreeNode nn;
        nn.cx = x; nn.cy = y; nn.px = node.cx; nn.py = node.cy;
        nn.heading = yaw; nn.velocity = vel;
        nn.g = 0;
        nn.g += calc_w(nn);             // LP:162 -- edge length only, not cumulative
        nn.cost = nn.g + calc_h(nn);    // LP:163
        out.push_back(nn);
      }
    }
    return out;
  }
  // LP:139-144
  static double calc_w(const TreeNode& n) {
    const double dx = n.cx - n.px, dy = n.cy - n.py;
    return std::sqrt(M::sq(dx) + M::sq(dy));
  }
  // LP:301-316 -- called as (child, parent): points run from the child AWAY from the parent
  std::vector<std::pair<double, double>> interpolate_points(double sx, double sy, double ex,
                                                            double ey) const {
    std::vector<std::pair<double, double>> pts;
    const double dx = sx - ex, dy = sy - ey;
    const double angle = M::atan2(dy, dx);
    const double dist = std::sqrt(M::sq(dx) + M::sq(dy));
    int n;
    if (!trunc_to_int(dist / P.search_res * 2, &n)) n = 0;
    for (int i = 0; i < n; ++i) {
      const double x = sx + i / static_cast<double>(n) * dist * M::cos(angle);
      const double y = sy + i / static_cast<double>(n) * dist * M::sin(angle);
      pts.emplace_back(x, y);
    }
    return pts;
  }
  // LP:269-287 (+ LP:289-299).  Unqualified abs() resolves to the double overload with
  // this toolchain; ORACLE_INT_ABS selects the int(abs(int)) reading.
  bool check_for_goal(const TreeNode& node, TreeNode* goal_node) const {
    const auto pts = interpolate_points(node.cx, node.cy, node.px, node.py);
    for (const auto& pt : pts) {
      // A18: pp_params.goal_test_int_abs (or the ORACLE_INT_ABS build switch) selects the int abs(int)
      // reading: the double differences are truncated toward zero by the implicit conversion
#ifdef ORACLE_INT_ABS
      const bool int_abs = true;
#else
      const bool int_abs = P.goal_test_int_abs != 0;
#endif
      const bool hit = int_abs
          ? (int_abs_within(pt.first - Q.goal_x, P.goal_pos_tolerance) &&
             int_abs_within(pt.second - Q.goal_y, P.goal_pos_tolerance) &&
             int_abs_within(node.velocity - Q.goal_speed, P.goal_speed_tolerance))
          : (std::fabs(pt.first - Q.goal_x) <= P.goal_pos_tolerance &&
             std::fabs(pt.second - Q.goal_y) <= P.goal_pos_tolerance &&
             std::fabs(node.velocity - Q.goal_speed) <= P.goal_speed_tolerance);
      if (hit) {
        const double dx = pt.first - node.cx, dy = pt.second - node.cy;  // LP:293-295
        TreeNode gn;
        gn.cx = pt.first; gn.cy = pt.second; gn.px = node.cx; gn.py = node.cy;
        gn.heading = M::atan2(dy, dx);
        gn.velocity = Q.goal_speed;
        gn.g = 0;
        gn.cost = 0;  // never assigned in the reference; the search returns at once
        *goal_node = gn;
        return true;
      }
    }
    return false;
  }

 private:
  // LP:84-96
  void initialize() {
    open_.clear();
    closed_.clear();
    frontier_ = decltype(frontier_)();
    if (G) O->visited.assign(static_cast<size_t>(G->width) * G->height, 0);
    TreeNode s;
    s.cx = 0; s.cy = 0; s.px = 0; s.py = 0;
    s.heading = 0;
    s.velocity = Q.start_speed;
    s.g = 0;
    s.cost = 9999999999.0;
    open_node(s);
  }
  // LP:98-102
  void open_node(TreeNode n) {
    n.id = static_cast<int>(O->slots.size());
    O->slots.push_back(n);
    O->closed_seq.push_back(-1);
    frontier_.push(n);
    open_.push_back(n);
  }
  // LP:104-113; returns the slot id that moved to the closed list
  int close_top() {
    const TreeNode& top = frontier_.top();
    const int pos = find_in(open_, top);
    int id = -1;
    if (pos >= 0) {
      id = open_[pos].id;
      O->closed_seq[id] = static_cast<int>(closed_.size());
      closed_.push_back(open_[pos]);
      open_.erase(open_.begin() + pos);
    }  // pos < 0 would be erase(end()) = UB in the reference; cannot happen
    frontier_.pop();
    return id;
  }
  static int find_in(const std::vector<TreeNode>& v, const TreeNode& n) {
    for (size_t k = 0; k < v.size(); ++k)
      if (same_node(v[k], n)) return static_cast<int>(k);
    return -1;
  }
  // LP:318-326 with a bounds guard (the reference has none)
  void mark_visited(double px, double py) {
    if (!G) return;
    int x, y;
    if (!trunc_to_int(G->width / 2 - py / G->res, &x)) return;
    if (!trunc_to_int(G->height / 2 - px / G->res, &y)) return;
    if (x < 0 || x >= G->width || y < 0 || y >= G->height) return;
    O->visited[static_cast<size_t>(x) * G->height + y] = 1;
  }
  // LP:503-527: walk parents by POINT equality, open list then closed list per pass,
  // following every link met during a pass (no break)
  bool calc_path(double gx, double gy) {
    std::vector<std::pair<double, double>> rev;
    double cx = gx, cy = gy;
    rev.emplace_back(cx, cy);
    const size_t guard = O->slots.size() + 2;
    while (!(cx == 0 && cy == 0)) {
      bool progressed = false;
      for (const TreeNode& n : open_)
        if (cx == n.cx && cy == n.cy) { cx = n.px; cy = n.py; rev.emplace_back(cx, cy); progressed = true; }
      for (const TreeNode& n : closed_)
        if (cx == n.cx && cy == n.cy) { cx = n.px; cy = n.py; rev.emplace_back(cx, cy); progressed = true; }
      if (!progressed || rev.size() > guard) return false;
    }
    O->path.assign(rev.rbegin(), rev.rend());  // LP:530-538
    return true;
  }
  // LP:542-583.  The loop bound `poses.size() - 2` is unsigned in the reference; fewer than
  // two poses is defined here as "no profile entries".
  void calc_profile() {
    const auto& path = O->path;
    if (path.size() < 2) return;
    for (size_t i = 0; i + 2 < path.size(); ++i) {
      const double ax = path[i].first, ay = path[i].second;
      const double bx = path[i + 1].first, by = path[i + 1].second;
      double cur_heading = 0, next_heading = 0, next_velocity = 0;  // GraphNode(cur, next, 0, 0)
      auto scan = [&](const std::vector<TreeNode>& v) {
        for (const TreeNode& n : v) {
          if (ax == n.cx && ay == n.cy) {
            cur_heading = n.heading;
          } else if (bx == n.cx && by == n.cy) {
            next_heading = n.heading;
            next_velocity = n.velocity;
          }
        }
      };
      scan(open_);
      scan(closed_);
      const double d_yaw = next_heading - cur_heading;
      O->yaw_rates.push_back(d_yaw / (P.time_step_ms / 1000));
      O->durations.push_back(P.time_step_ms);
      O->speeds.push_back(next_velocity);
    }
  }
  void finish(int status) {
    O->status = status;
    O->n_open = static_cast<int>(open_.size());
    O->n_closed = static_cast<int>(closed_.size());
    // final state of every slot (LP:68 may have overwritten open-list copies)
    for (const TreeNode& n : open_) O->slots[n.id] = n;
    for (const TreeNode& n : closed_) O->slots[n.id] = n;
  }

  pp_params P;
  pp_query Q{};
  const CSpace* G;
  Lattice lattice_;
  PlanOutput* O = nullptr;
  std::priority_queue<TreeNode, std::vector<TreeNode>, CheaperCost> frontier_;  // local_planner.hpp:111
  std::vector<TreeNode> open_, closed_;                                         // local_planner.hpp:112-113
};

// "first node in scan order (open list, then closed list) whose child_point == (x, y)"
// -- the lookup rule of LP:511-526, used to define parent ids and path node ids
inline int first_node_with_child(const PlanOutput& o, double x, double y) {
  int best_open = -1;
  int best_closed = -1, best_closed_seq = 0;
  for (size_t id = 0; id < o.slots.size(); ++id) {
    const TreeNode& n = o.slots[id];
    if (!(n.cx == x && n.cy == y)) continue;
    if (o.closed_seq[id] < 0) {
      if (best_open < 0) best_open = static_cast<int>(id);
    } else if (best_closed < 0 || o.closed_seq[id] < best_closed_seq) {
      best_closed = static_cast<int>(id);
      best_closed_seq = o.closed_seq[id];
    }
  }
  return best_open >= 0 ? best_open : best_closed;
}

}  // namespace oracle
