// Runs ros/local_planner_b200_node.cpp -- the drop-in replacement of the reference executable
// local_planner_node -- UNCHANGED against the ROS test doubles of oracle/ros_doubles/ (the same
// doubles the reference's own node is compiled against for oracle/_ref), on a GPU box.
// A scenario file scripts what ROS would deliver; everything the node publishes is printed with
// %a (exact) so that tests/test_gpu_ros_shim.py can compare it with what the REFERENCE node
// publishes for the same messages.
//
// scenario lines:  param <name> <value> | tf <x> <y> <yaw> | grid <file> <W> <H> <res>
//                  twist <x> <y> <z> | nav <x> <y> <speed> <max_speed> <max_accel>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#define main shim_main
#include "../../ros/local_planner_b200_node.cpp"
#undef main

namespace {

struct Cmd { std::string op; std::vector<std::string> a; };

void print_outputs() {
  auto& goals = ros_doubles::captured<geometry_msgs::PoseStamped>();
  auto& paths = ros_doubles::captured<nav_msgs::Path>();
  auto& mps = ros_doubles::captured<prius_msgs::MotionPlanning>();
  std::printf("goal %zu", goals.size());
  for (auto& g : goals) std::printf(" %a %a %s", g.pose.position.x, g.pose.position.y, g.header.frame_id.c_str());
  std::printf("\npath %zu", paths.size());
  if (!paths.empty()) {
    std::printf(" %s %zu", paths.back().header.frame_id.c_str(), paths.back().poses.size());
    for (auto& p : paths.back().poses) std::printf(" %a %a", p.pose.position.x, p.pose.position.y);
  }
  std::printf("\nmp %zu", mps.size());
  if (!mps.empty()) {
    const auto& m = mps.back();
    std::printf(" %a %a %zu", m.max_speed, m.max_accel, m.yaw_rates.size());
    for (size_t k = 0; k < m.yaw_rates.size(); ++k) std::printf(" %a %a %a", m.yaw_rates[k], m.durations[k], m.speeds[k]);
  }
  std::printf("\n");
  goals.clear(); paths.clear(); mps.clear();
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s scenario.txt\n", argv[0]); return 2; }
  std::ifstream in(argv[1]);
  std::vector<Cmd> script;
  for (std::string line; std::getline(in, line);) {
    std::istringstream ls(line);
    Cmd c;
    if (!(ls >> c.op)) continue;
    for (std::string w; ls >> w;) c.a.push_back(w);
    script.push_back(c);
  }
  // parameters and tf must exist before the node's constructor runs (LP:14-17 reads them there)
  for (const Cmd& c : script) {
    if (c.op == "param") ros_doubles::state().params[c.a[0]] = std::stod(c.a[1]);
    if (c.op == "tf") {
      tf::Quaternion q; q.setRPY(0, 0, std::stod(c.a[2]));
      tf::register_transform("base_link", "map", tf::Transform(q, tf::Vector3(std::stod(c.a[0]), std::stod(c.a[1]), 0)));
    }
  }
  ros_doubles::state().quiet = false;   // ROS_ERROR_STREAM of the node -> stderr
  ros_doubles::state().spin_script = [&]() {
    for (const Cmd& c : script) {
      if (c.op == "tf") {
        tf::Quaternion q; q.setRPY(0, 0, std::stod(c.a[2]));
        tf::register_transform("base_link", "map", tf::Transform(q, tf::Vector3(std::stod(c.a[0]), std::stod(c.a[1]), 0)));
      } else if (c.op == "grid") {
        nav_msgs::OccupancyGrid g;
        g.info.width = std::stoi(c.a[1]); g.info.height = std::stoi(c.a[2]); g.info.resolution = std::stof(c.a[3]);
        g.data.resize(static_cast<size_t>(g.info.width) * g.info.height);
        std::ifstream gf(c.a[0], std::ios::binary);
        gf.read(reinterpret_cast<char*>(g.data.data()), g.data.size());
        ros_doubles::deliver("/local_costmap", g);
      } else if (c.op == "twist") {
        gazebo_msgs::ModelStates ms;
        ms.name = {"ground_plane", "prius"};
        ms.pose.resize(2); ms.twist.resize(2);
        ms.twist[1].linear.x = std::stod(c.a[0]); ms.twist[1].linear.y = std::stod(c.a[1]); ms.twist[1].linear.z = std::stod(c.a[2]);
        ros_doubles::deliver("/gazebo/model_states", ms);
      } else if (c.op == "nav") {
        prius_msgs::LocalNav nav;
        nav.x = std::stod(c.a[0]); nav.y = std::stod(c.a[1]); nav.speed = std::stod(c.a[2]);
        nav.max_speed = std::stod(c.a[3]); nav.max_accel = std::stod(c.a[4]);
        const int n = ros_doubles::deliver("/local_nav_waypoints", nav);
        std::printf("delivered %d\n", n);
        print_outputs();
      }
    }
  };
  char* args[] = {argv[0], nullptr};
  return shim_main(1, args);
}
